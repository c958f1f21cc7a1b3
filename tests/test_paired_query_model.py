"""Host model of the opt-in paired query kernel's candidate logic (csrc/collide_paired.cu, MSIM_QUERY_PAIRED=1): two adjacent slots of the
cell order per thread; when they lie in the same row and in the same or in neighbouring cells both are tested against the UNION of their
candidate runs without any per-entity range test (a candidate outside an entity's own 3 x 2 cells is more than a radius away).  The model walks
the slot pairs exactly as the kernel does - run bounds from the prefix table, union / separate paths, the pair (slot 0, slot 1) itself, the
look-above fallback - with the device's binary32 arithmetic, and must reproduce the oracle's unique-pair count and colour flags.  The kernel
itself has not run on hardware yet (tests/test_zz_gpu_unverified.py); this pins the algorithm it implements."""
import numpy as np
import pytest

from conftest import to_oracle_entities

f32 = np.float32


def run(M, O, m, n, radius, passes, seed):
    om = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)
    e = to_oracle_entities(O, m.init_entities(n, seed=seed))
    for _ in range(passes): O.move_pass(e, om, threads=8)
    want_pairs = O.collide_pass(e, om.world_w, om.world_h, radius, threads=8)
    want_flags = O.collision_flags(e)
    g = M.grid_params(m.width, m.height, radius); inv, T, ncx, ncy = g["inv_cell"], g["hit_threshold"], g["cells_x"], g["cells_y"]
    pos = e["pos"].astype(np.float32)
    cx = np.clip(np.floor(pos[:,0]*inv),0,ncx-1).astype(np.int64); cy = np.clip(np.floor(pos[:,1]*inv),0,ncy-1).astype(np.int64)
    key = cy*ncx+cx
    order = np.argsort(key, kind="stable")
    sp = pos[order]; skey = key[order]; scx=cx[order]; scy=cy[order]
    cell_start = np.searchsorted(skey, np.arange(ncx*ncy+1))
    def runs(j):
        x0=max(scx[j]-1,0); x1=min(scx[j]+1,ncx-1)
        own_lo=cell_start[scy[j]*ncx+x0]; own_hi=cell_start[scy[j]*ncx+x1+1]
        if scy[j]>0: ab_lo=cell_start[(scy[j]-1)*ncx+x0]; ab_hi=cell_start[(scy[j]-1)*ncx+x1+1]
        else: ab_lo=ab_hi=0
        return own_lo,own_hi,ab_lo,ab_hi,x0,x1
    def d2(q,p):
        dx=(q[...,0]-p[0]).astype(f32); dy=(q[...,1]-p[1]).astype(f32)
        return ((dx*dx).astype(f32)+(dy*dy).astype(f32)).astype(f32)
    def cnt(a,b,p): return int((d2(sp[a:b],p)<T).sum()) if b>a else 0
    def anyr(a,b,p): return bool((d2(sp[a:b],p)<T).any()) if b>a else False
    N=n; total=0; flags=np.zeros(N,dtype=np.uint8); together_n=0
    for j0 in range(0,N,2):
        j1=j0+1; v1=j1<N
        r0=runs(j0); r1=runs(j1) if v1 else None
        c0=c1=0
        tog = v1 and scy[j0]==scy[j1] and 0 <= scx[j1]-scx[j0] <= 1
        if tog:
            together_n+=1
            if r0[2]<r1[3]:
                c0+=cnt(r0[2],r1[3],sp[j0]); c1+=cnt(r0[2],r1[3],sp[j1])
            c0+=cnt(r0[0],j0,sp[j0]); c1+=cnt(r0[0],j0,sp[j1])
            c1+= int(d2(sp[j0],sp[j1])<T)
        else:
            if r0[2]<r0[3]: c0+=cnt(r0[2],r0[3],sp[j0])
            c0+=cnt(r0[0],j0,sp[j0])
            if v1:
                if r1[2]<r1[3]: c1+=cnt(r1[2],r1[3],sp[j1])
                c1+=cnt(r1[0],j1,sp[j1])
        total+=c0+c1
        for (j,c,r) in ((j0,c0,r0),(j1,c1,r1)):
            if r is None: continue
            hit=c!=0
            if not hit:
                hit=anyr(max(j+1,r[0]),r[1],sp[j])
                if not hit and scy[j]+1<ncy:
                    lo=cell_start[(scy[j]+1)*ncx+r[4]]; hi=cell_start[(scy[j]+1)*ncx+r[5]+1]
                    hit=anyr(lo,hi,sp[j])
            flags[j]=hit
    got_flags=np.zeros(N,dtype=np.uint8); got_flags[order]=flags
    assert total==want_pairs and (got_flags==want_flags).all()
    return together_n, (N + 1) // 2


@pytest.mark.parametrize("case", ["city_r10", "city_odd_r3.3", "city_r25", "test_map_dense", "lattice", "tiny"])
def test_paired_query_model_reproduces_the_oracle(msim, orc, small_city, test_map, case):
    M, O = msim, orc
    if case == "city_r10":
        together, pairs = run(M, O, small_city, 6000, 10.0, 120, 1)
        assert together > pairs // 3  # the union path is common even at this low density (75 % at 20 k entities)
    elif case == "city_odd_r3.3":
        run(M, O, small_city, 5001, 3.3, 40, 2)
    elif case == "city_r25":
        run(M, O, small_city, 2000, 25.0, 300, 3)
    elif case == "test_map_dense":
        together, pairs = run(M, O, test_map, 999, 10.0, 77, 4)  # 124 k pairs among 999 entities on four roads
        assert together >= pairs - 2
    elif case == "lattice":
        run(M, O, M.Map.grid(24, 17, 20.0), 3000, 10.0, 50, 5)
    else:
        run(M, O, small_city, 7, 10.0, 3, 6)
