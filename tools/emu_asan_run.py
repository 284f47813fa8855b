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
    if a>256:
        r=emu.fill_pk(t,s,1,1,1,K=4,R=2,grid=1,split=1,count=8,hx=True)
        assert r['count']==o.count
tops=[bytes(random.choice(b"ACGT") for _ in range(n)) for n in (256,300,17,700)]
sides=[bytes(random.choice(b"ACGT") for _ in range(n)) for n in (256,40,130,90)]
r=emu.fill_batch(tops,sides,1,1,1,grid=1)
print("asan run ok")
