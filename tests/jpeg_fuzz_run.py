"""Helper of tests/test_jpeg_core.py: mutated JPEG files (header bytes flipped, truncated, entropy bytes flipped, garbage) through
jd::parse and the decoder's logic (host simulation with the kernel's own symbol loop), built with AddressSanitizer: whatever a
camera or a corrupted log delivers must end in an error code or in some image, never in an out-of-bounds access.
usage: jpeg_fuzz_run.py <libjpc_asan.so> <seed> <count>"""
import ctypes as C, numpy as np, cv2, sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import realset
from oracle import synth
lib=C.CDLL(sys.argv[1])
def run(buf, shape, fn='jpc_decode_gpu_loop'):
    b=np.ascontiguousarray(buf,np.uint8); r=C.c_int(); n=C.c_int(); W=C.c_int(); H=C.c_int()
    rc=lib.jpc_info(b.ctypes.data_as(C.c_void_p),C.c_size_t(len(b)),C.byref(W),C.byref(H))
    if rc: return rc
    if W.value*H.value>40_000_000: return -99
    out=np.zeros((H.value,W.value,3),np.uint8)
    return getattr(lib,fn)(b.ctypes.data_as(C.c_void_p),C.c_size_t(len(b)),out.ctypes.data_as(C.c_void_p),0,C.byref(r),C.byref(n))
rng=np.random.default_rng(int(sys.argv[2]))
bases=[np.asarray(realset.jpeg(0)).copy(), cv2.imencode('.jpg',synth.frame(1,123,161),[cv2.IMWRITE_JPEG_QUALITY,80])[1].ravel().copy(),
       cv2.imencode('.jpg',cv2.cvtColor(synth.frame(1,64,96),cv2.COLOR_BGR2GRAY))[1].ravel().copy()]
shapes=[(480,640,3),(123,161,3),(64,96,3)]
codes={}
N=int(sys.argv[3])
for it in range(N):
    k=it%3; b=bases[k].copy(); mode=rng.integers(0,4)
    if mode==0:   # flip bytes in the header (first 700 bytes)
        for _ in range(rng.integers(1,6)): b[rng.integers(2,min(700,len(b)))]=rng.integers(0,256)
    elif mode==1: # truncate
        b=b[:rng.integers(4,len(b))]
    elif mode==2: # flip bytes in entropy data
        for _ in range(rng.integers(1,20)): b[rng.integers(600,len(b))]=rng.integers(0,256)
    else:         # random garbage after SOI
        b=np.concatenate([b[:2], rng.integers(0,256,rng.integers(1,400),dtype=np.uint8)])
    rc=run(b,shapes[k]); codes[rc]=codes.get(rc,0)+1
print('fuzz ok', codes)
