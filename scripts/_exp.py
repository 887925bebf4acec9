import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uc2_b200 import _lib
dev="cuda"; bf=torch.bfloat16
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
H,F=768,3072
Mfull=19200
g=(torch.randn(Mfull,F,device=dev)*0.05).to(bf); w2=(torch.randn(H,F,device=dev)*0.05).to(bf)
res=torch.randn(Mfull,H,device=dev); z=torch.empty(Mfull,H,device=dev); bias=torch.zeros(H,device=dev)
for M,c,b in [(19200,2,256),(18944,2,256),(18944+256*3,2,256),(256,2,256),(256,1,64),(256,1,128),(256,2,128),(512,1,64),(19200,0,0)]:
    us=t(lambda: _lib.gemm(g,w2,M,H,F,bias=bias,residual=res,out_f32=z,block_n=b,ctas=c))
    print(f"FFN2 fwd M={M} ctas={c} bn={b}: {us:.1f} us")
x=(torch.randn(Mfull,H,device=dev)*0.05).to(bf); wq=(torch.randn(2304,H,device=dev)*0.05).to(bf); o=torch.empty(Mfull,2304,dtype=bf,device=dev); bq=torch.zeros(2304,device=dev)
for M,c,b in [(19200,2,256),(18944,2,256),(256,1,64),(256,2,128),(256,2,256),(19200,0,0)]:
    us=t(lambda: _lib.gemm(x,wq,M,2304,H,bias=bq,out_bf16=o,block_n=b,ctas=c))
    print(f"QKV fwd M={M} ctas={c} bn={b}: {us:.1f} us")
# empty-ish kernel launch cost: tiny gemm
a=torch.zeros(128,64,dtype=bf,device=dev); bb=torch.zeros(64,64,dtype=bf,device=dev); oo=torch.empty(128,64,dtype=bf,device=dev)
print("tiny gemm c1/64:", t(lambda: _lib.gemm(a,bb,128,64,64,out_bf16=oo,block_n=64,ctas=1)))
a=torch.zeros(256,64,dtype=bf,device=dev); bb=torch.zeros(128,64,dtype=bf,device=dev); oo=torch.empty(256,128,dtype=bf,device=dev)
print("tiny gemm c2/128:", t(lambda: _lib.gemm(a,bb,256,128,64,out_bf16=oo,block_n=128,ctas=2)))
