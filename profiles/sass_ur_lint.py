"""Lint of the shipped SASS for one hazard: a UNIFORM register that holds different values on different paths of a divergent region.

Uniform registers (UR0..UR63, UP0..UP7) exist once per warp.  Lanes of a warp that sit on different paths of a divergent region (between a
BSSY and the BSYNC it names) are interleaved by the hardware, so if one path loads the row count into UR8 and another path the distance
threshold, a lane can compare against the other path's number.  Round 2 met exactly that in the collision query (profiles/
r2_flag_count_race.md: one missed neighbour per ~10^10 entity-ticks, one missed count per ~10 ticks).  ptxas does it silently; this lint reads
`cuobjdump -sass` of the built library and reports, for every outermost BSSY region of every kernel, uniform registers that

  * are written inside the region by two instructions with different text (two different values), or
  * are written inside the region after its first branch AND read inside the region by an instruction that no write in the region
    dominates in address order (a value from outside the region is live on some path while another path overwrites it).

Writes that every lane executes before the region's first branch are harmless (the warp is still together), and so are identical writes
(same opcode and operands: every path loads the same value).  The rule is conservative in address order, not a control-flow analysis; a
finding is a reason to read the region, not a proof."""
import re
import subprocess
import sys

INSN = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")
UREG = re.compile(r"\bU(?:R\d+|P\d+)\b")
# opcodes whose FIRST operand(s) are uniform destinations
UDEST_OPS = ("LDCU", "UMOV", "S2UR", "UIADD3", "ULEA", "USHF", "ULOP3", "UISETP", "UIMAD", "USEL", "UFLO", "UPOPC", "UBREV", "UBMSK", "UPRMT",
             "R2UR", "REDUX", "VOTEU", "UCLEA", "UF2FP", "UPLOP3", "UP2UR", "UR2UP", "UCGABAR", "USGXT", "UFADD", "UFMUL", "UFFMA", "UMNMX", "UVIADD",
             "UVIMNMX", "UI2F", "UF2I", "UI2FP", "UF2F", "UFSETP", "UFSEL", "UFMNMX", "ULEPC", "UGETNEXTWORKID")


def kernels(sass: str):
    name, body = None, []
    for line in sass.splitlines():
        if "Function :" in line:
            if name:
                yield name, body
            name, body = line.split("Function :")[1].strip(), []
            continue
        m = INSN.match(line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


def split_pred(text: str):
    return re.sub(r"^@!?U?P\d+\s+", "", text)


def udests(text: str):
    t = split_pred(text)
    head = t.split()[0]
    op = head.split(".")[0]
    if op not in UDEST_OPS:
        return []
    ops = [o.strip() for o in t[len(head):].split(",")]
    out = []
    m = UREG.fullmatch(ops[0]) if ops else None
    if m:
        out.append(ops[0])
    # a second destination is always a uniform PREDICATE (carry-out, second compare result, vote predicate)
    if len(ops) > 1 and re.fullmatch(r"UP\d+", ops[1]) and op in ("UISETP", "UIADD3", "ULEA", "VOTEU", "UPLOP3", "UFSETP"):
        out.append(ops[1])
    if (".64" in head or ".WIDE" in head) and out and out[0].startswith("UR"):
        out.append("UR%d" % (int(out[0][2:]) + 1))
    return [r for r in out if r not in ("URZ", "UPT")]


def uuses(text: str):
    t = split_pred(text)
    d = set(udests(text))
    regs = UREG.findall(t)
    # descriptor operands `desc[UR4][R2.64]` are 64-bit pairs
    for m in re.finditer(r"desc\[UR(\d+)\]", t):
        regs.append("UR%d" % (int(m.group(1)) + 1))
    first = True
    uses = []
    for r in regs:
        if r in d and first:
            first = False
            continue
        uses.append(r)
    g = re.match(r"^@!?(UP\d+)", text)
    if g:
        uses.append(g.group(1))
    return uses


def lint_kernel(body):
    findings = []
    i = 0
    while i < len(body):
        addr, text = body[i]
        m = re.match(r"(?:@!?U?P\d+\s+)?BSSY\S*\s+B\d+,\s*0x([0-9a-f]+)", text)
        if not m:
            i += 1
            continue
        end = int(m.group(1), 16)
        region = [(a, t) for a, t in body[i + 1:] if a < end]
        first_branch = next((k for k, (_, t) in enumerate(region) if re.search(r"\bBRA\b|\bBRX\b|\bJMP\b", t)), len(region))
        writes = {}
        for k, (a, t) in enumerate(region):
            for r in udests(t):
                writes.setdefault(r, []).append((k, a, split_pred(t)))
        for r, ws in writes.items():
            texts = {w[2] for w in ws}
            if len(texts) > 1:
                findings.append((addr, r, "written with different values inside one divergent region: " + " | ".join(f"{w[1]:04x} {w[2]}" for w in ws)))
                continue
            if all(w[0] < first_branch for w in ws):
                continue
            first_write = min(w[0] for w in ws)
            early_reads = [(a, t) for k, (a, t) in enumerate(region) if first_branch <= k < first_write and r in uuses(t)]
            if early_reads:
                findings.append((addr, r, f"read at {early_reads[0][0]:04x} ({early_reads[0][1]}) with a value from outside the region, overwritten at "
                                          f"{ws[0][1]:04x} ({ws[0][2]}) on another path"))
        i += 1 + len(region)  # outermost regions only: nested BSSY ranges are inside `region`
    return findings


def lint_library(path: str, only: str | None = None):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out = {}
    for name, body in kernels(sass):
        if only and only not in name:
            continue
        f = lint_kernel(body)
        if f:
            out[name] = f
    return out


if __name__ == "__main__":
    res = lint_library(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
    for k, fs in res.items():
        print(k)
        for addr, reg, why in fs:
            print(f"   region at {addr:04x}: {reg}: {why}")
    print(f"{sum(len(v) for v in res.values())} finding(s) in {len(res)} kernel(s)")
    sys.exit(1 if res else 0)
