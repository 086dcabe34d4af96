import torch, time
dev="cuda"
B,J,H=200000,500,64
x=(torch.randint(0,3,(B,J),device=dev)-1).float()
xb=x.bfloat16()
W=torch.randn(H,J,device=dev)*0.1
def split3(w):
    hi=w.bfloat16(); r=w-hi.float(); mid=r.bfloat16(); lo=(r-mid.float()).bfloat16()
    return hi,mid,lo
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
ref=(x.double()@W.double().t())
print("fp32 mm fwd ms", t(lambda: x@W.t()), "err", ((x@W.t()).double()-ref).abs().max().item()/ref.abs().max().item())
torch.backends.cuda.matmul.allow_tf32=False
W3=torch.cat(split3(W),0)   # [192, J]
try:
    f=lambda: torch.mm(xb, W3.t(), out_dtype=torch.float32)
    o=f(); pre=o[:,:H]+o[:,H:2*H]+o[:,2*H:]
    print("bf16x3 mm out_dtype fwd ms", t(f), "err", (pre.double()-ref).abs().max().item()/ref.abs().max().item())
except Exception as e:
    print("mm out_dtype failed:", repr(e)[:300])
# backward: dW = dpre^T X
dpre=torch.randn(B,H,device=dev)
refw=dpre.double().t()@x.double()
print("fp32 mm bwd ms", t(lambda: dpre.t()@x), "err", ((dpre.t()@x).double()-refw).abs().max().item()/refw.abs().max().item())
d3=torch.cat(split3(dpre),1)  # [B,192]
for C in (1,16,64,128):
    try:
        n=B//C
        def g():
            a=d3.view(C,n,3*H).transpose(1,2)       # [C,192,n]
            return torch.bmm(a, xb.view(C,n,J), out_dtype=torch.float32).sum(0)
        o=g(); dw=o[:H]+o[H:2*H]+o[2*H:]
        print("bf16x3 bmm C=",C," ms", t(g), "err", (dw.double()-refw).abs().max().item()/refw.abs().max().item())
    except Exception as e:
        print("bmm out_dtype failed:", repr(e)[:300]); break
print("split3 ms", t(lambda: torch.cat(split3(dpre),1)))
