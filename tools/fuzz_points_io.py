"""Randomised soak of the native savetxt / loadtxt against numpy (host only): random shapes, magnitudes over the whole double
range, random bit patterns (NaN, inf, denormals), float32 input, 1-D arrays, empty files.   python tools/fuzz_points_io.py"""
import sys, os, tempfile, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, warnings
from pope_b200 import points_io
rng=np.random.default_rng(0)
d=tempfile.mkdtemp()
bad=0
for case in range(400):
    rows=int(rng.integers(0,30)); cols=int(rng.integers(1,6))
    kind=rng.integers(0,4)
    if kind==0: a=rng.standard_normal((rows,cols))
    elif kind==1: a=(rng.standard_normal((rows,cols))*10.0**rng.integers(-300,300,(rows,cols)))
    elif kind==2: a=np.frombuffer(rng.bytes(rows*cols*8),dtype=np.float64).reshape(rows,cols).copy()   # random bit patterns (nan, inf, denormals)
    else: a=rng.standard_normal((rows,cols)).astype(np.float32)
    if rng.random()<0.3 and rows: a=a[:,0].copy()
    p1,p2=os.path.join(d,"a.txt"),os.path.join(d,"b.txt")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.savetxt(p1,a); points_io.savetxt(p2,a)
        if open(p1,'rb').read()!=open(p2,'rb').read(): bad+=1; print("savetxt differs",case,a.shape,a.dtype)
        want=np.loadtxt(p1,delimiter=' ') if a.size else np.loadtxt(p1)
        got=points_io.loadtxt(p1)
    if got.shape!=want.shape or not np.array_equal(got,want,equal_nan=True) or not np.array_equal(np.signbit(got[~np.isnan(got)]),np.signbit(want[~np.isnan(want)])):
        bad+=1; print("loadtxt differs",case,a.shape,got.shape,want.shape)
try:
    import cv2
    for case in range(300):
        h,w=int(rng.integers(1,70)),int(rng.integers(1,70)); c=int(rng.choice([1,3,4]))
        a=rng.integers(0,256,(h,w,c),dtype=np.uint8) if rng.random()<0.5 else np.repeat(rng.integers(0,256,(h,1,c),dtype=np.uint8),w,1)
        if c==1 and rng.random()<0.5: a=a[:,:,0].copy()
        p1,p2=os.path.join(d,"a.png"),os.path.join(d,"b.png")
        points_io.imwrite_png(p1,a,level=int(rng.integers(0,10))); cv2.imwrite(p2,a)
        x,y=cv2.imread(p1,cv2.IMREAD_UNCHANGED),cv2.imread(p2,cv2.IMREAD_UNCHANGED)
        if x is None or x.shape!=y.shape or not np.array_equal(x,y): bad+=1; print("png differs",case,a.shape)
    print("png: 300 cases")
except ImportError:
    print("cv2 not importable: PNG cases skipped")
print("points_io fuzz: 400 cases,",bad,"failures")
