import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raiko_b200 as rk
from raiko_b200 import _native
lib = _native.load()
n = int(sys.argv[1]); wb = int(sys.argv[2]); host = len(sys.argv) > 3 and sys.argv[3] == "host"
s = rk.KzgSettings(window_bits=wb)
g = torch.Generator(device="cuda"); g.manual_seed(1)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device="cuda", generator=g)
blobs[:, :, 0] %= 0x73
dev = "cpu" if host else "cuda"
if host: blobs = blobs.cpu()
outs = {k: torch.zeros((n, w), dtype=torch.uint8, device=dev) for k, w in (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))}
st = lib.rk_commit_batch(s._ctx, blobs.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["st"].data_ptr())
print("commit", st, _native.last_error() if st else "", flush=True)
st = lib.rk_commit_prove_batch(s._ctx, blobs.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["x"].data_ptr(), outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
print("commit_prove", st, _native.last_error() if st else "", flush=True)
torch.cuda.synchronize()
print("ok", outs["c"][0, :8].cpu().numpy().tobytes().hex(), int(outs["st"].sum()))
