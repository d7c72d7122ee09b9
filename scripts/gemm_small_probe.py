"""Where does the time of a SHORT GEMM go?  Graph-timed (no host overhead) sweeps of K, tile shape and the
VLK_GEMM_DEBUG bring-up switches (1 = no epilogue work, 2 = no TMA loads / full-barrier waits), next to cuBLAS."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import ops

BF = torch.bfloat16


def graph_time(fn, iters=20):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * iters) * 1e3


def case(M, N, K, res=False, tag=""):
    a = torch.randn(M, K, device="cuda").to(BF); w = torch.randn(N, K, device="cuda").to(BF)
    bias = torch.randn(N, device="cuda").to(BF); r = torch.randn(M, N, device="cuda").to(BF)
    out = torch.empty(M, N, device="cuda", dtype=BF)
    kw = dict(bias=bias, residual=r) if res else {}
    us = graph_time(lambda: ops.gemm(a, w, out=out, **kw))
    cb = graph_time(lambda: torch.matmul(a, w.t(), out=out))
    line = f"M={M:6d} N={N:5d} K={K:5d} res={int(res)} {tag:22s}: vlk {us:7.1f} us ({2.0*M*N*K/us/1e6:6.0f} TF)  cuBLAS {cb:7.1f} us"
    for dbg in (1, 2, 3):
        os.environ["VLK_GEMM_DEBUG"] = str(dbg)
        line += f"  dbg{dbg} {graph_time(lambda: ops.gemm(a, w, out=out, **kw)):6.1f}"
    os.environ.pop("VLK_GEMM_DEBUG")
    print(line, flush=True)


for K in (64, 256, 768, 1536, 3072):
    case(4096, 768, K)
case(4096, 768, 768, res=True)
case(256, 256, 768, tag="one tile")
case(256, 256, 64, tag="one tile, one k-block")
for bn, cl in ((256, 1), (128, 1), (64, 1), (128, 3)):
    os.environ["VLK_GEMM_BN"], os.environ["VLK_GEMM_CLUSTER"] = str(bn), str(cl)
    case(4096, 768, 768, tag=f"bn={bn} cluster={cl}")
    case(4096, 2304, 768, tag=f"bn={bn} cluster={cl}")
    case(4096, 768, 3072, tag=f"bn={bn} cluster={cl}")
os.environ.pop("VLK_GEMM_BN"); os.environ.pop("VLK_GEMM_CLUSTER")
case(16448, 1024, 1024, res=True)
case(16448, 1024, 1024, res=False)
case(16384, 768, 768, res=True)
case(16384, 3072, 768)
