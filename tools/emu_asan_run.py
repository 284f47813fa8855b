import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import emu, oracle, numpy as np, random
emu._lib = None
real_build = emu.build
emu.build = lambda: None
import os
emu.HERE_SAVE = emu.HERE
orig = C.CDLL
def patched(path, *a, **k):
    if path.endswith("libnwb_emu.so"): path = "/tmp/libnwb_emu_asan.so"
    return orig(path, *a, **k)
emu.C.CDLL = patched
random.seed(4)
for (a,b) in [(7,7),(300,100),(513,70),(600,260),(90,700)]:
    t=bytes(random.choice(b"ACGT") for _ in range(a)); s=bytes(random.choice(b"ACGT") for _ in range(b))
    o=oracle.fill(t,s,1,1,1)
    for R in (1,2):
        for cnt in (False,True):
            r=emu.fill_pk(t,s,1,1,1,K=4,R=R,grid=2,count=cnt)
            assert r['opt_score']==o.final_score
    r=emu.fill_i32(t,s,1,1,1,flags=1|2|8|0x20,grid=2)
    assert r['opt_score']==o.final_score
    r=emu.fill_pk(t,s,1,1,1,K=4,R=2,grid=2,split=1,count=True) if a>256 else None
    # sweeping + flush warp kernel (ring, flags) and the count sweep over its arrow codes (staged ring for 8 cells per lane)
    for cpl in (0,2,4,8):
        r=emu.fill_pk(t,s,1,1,1,K=4,R=2,grid=2,count=cpl,hx=True)
        assert r['opt_score']==o.final_score and (cpl==0 or r['count']==o.count)
    # queue mode of the same kernel: ticketed blocks, 1 / 2 / 3 adjacent strips per block, the acknowledgement check
    # of a pipelined strip group in front of the last local strip's remote stores (split)
    for hxq in (4,5,6):
        r=emu.fill_pk(t,s,1,1,1,K=4,R=2,grid=2,hx=hxq,split=(1 if a>256 else 0))
        assert r['opt_score']==o.final_score and r['branch_count']==o.branch_count
    # the three-rows-per-lane geometry (nwb_fill_hy.cuh): stream words from step 0 on, 47 groups early; per-lane row parity
    for cpl in (0,8):
        r=emu.fill_pk(t,s,2,1,2,K=4,R=2,grid=2,count=cpl,hx=2,split=(1 if a>256 else 0))
        o2=oracle.fill(t,s,2,1,2)
        assert r['opt_score']==o2.final_score and r['branch_count']==o2.branch_count and (cpl==0 or r['count']==o2.count)
    if a>256:
        r=emu.fill_pk(t,s,1,1,1,K=4,R=2,grid=1,split=1,count=8,hx=True)
        assert r['count']==o.count
tops=[bytes(random.choice(b"ACGT") for _ in range(n)) for n in (256,300,17,700)]
sides=[bytes(random.choice(b"ACGT") for _ in range(n)) for n in (256,40,130,90)]
r=emu.fill_batch(tops,sides,1,1,1,grid=1,count=True)
assert r['kernel']=='pk'
for i in range(4):
    assert int(r['counts'][i])==oracle.fill(tops[i],sides[i],1,1,1).count
# two pairs per warp (ragged partners, odd batch, table edges at the pitch boundary) and the batch count pass
lens=[(256,256),(1,1),(255,257),(3,40),(0,5),(256,31),(100,300),(5,0),(250,64),(7,65),(256,2)]
tops=[bytes(random.choice(b"ACGT") for _ in range(a)) for a,_ in lens]
sides=[bytes(random.choice(b"ACGT") for _ in range(b)) for _,b in lens]
r=emu.fill_batch(tops,sides,2,1,2,grid=1,bx=1,count=True)
for i in range(len(lens)):
    o=oracle.fill(tops[i],sides[i],2,1,2)
    assert (r['scores'][i],r['branches'][i],int(r['counts'][i]))==(o.final_score,o.branch_count,o.count), i
# uniform shapes swept back to back, 12 and 16 warps per block (64- and 48-row rings)
for (a,b,n) in ((256,256,3),(40,64,61),(255,96,27)):
    tops=[bytes(random.choice(b"ACGT") for _ in range(a)) for _ in range(n)]
    sides=[bytes(random.choice(b"ACGT") for _ in range(b)) for _ in range(n)]
    for w in ("12","16"):
        os.environ["NWB_CX_WARPS"]=w
        r=emu.fill_batch(tops,sides,1,1,1,grid=1,bx=2,count=True)
        assert r['kernel']=='cx'
        for i in range(n):
            o=oracle.fill(tops[i],sides[i],1,1,1)
            assert (r['scores'][i],r['branches'][i],int(r['counts'][i]))==(o.final_score,o.branch_count,o.count), i
    del os.environ["NWB_CX_WARPS"]
# bit-parallel batch kernel (one thread per pair: match vectors, look-up table, staging buffer in shared memory; word
# loads at the end of the string buffers; the left-over list) and the per-lane sparse count (16-byte loads around the
# window, the dense kernel over its left-over list)
lens=[(256,256)]*33+[(200,90),(256,31),(17,130),(1,1),(64,64),(0,3),(5,0),(255,77),(129,300),(3,3),(256,1)]*3
tops=[bytes(random.choice(b"ACGTNR" if i%13==5 else (b"ACGTN" if i%13==6 else b"ACGT")) for _ in range(a)) for i,(a,_) in enumerate(lens)]
sides=[bytes(random.choice(b"ACGTX") for _ in range(b)) for _,b in lens]
r=emu.fill_batch_bp(tops,sides,1,1,1,grid=2,warps=2)
assert r is not None and r['n_fallback']>=1
counts,nfb=emu.batch_lcount(tops,sides,r['tables'],grid=2,warps=2)
for i in range(len(lens)):
    o=oracle.fill(tops[i],sides[i],1,1,1)
    assert (r['scores'][i],r['branches'][i],int(counts[i]))==(o.final_score,o.branch_count,o.count), i
print("asan run ok")
