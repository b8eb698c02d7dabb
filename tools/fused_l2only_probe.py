"""What would the fused OPNet forward cost per frame WITHOUT the LSTM1 / who-to-track work on its SMs?  Runs the shipped kernel,
then the experiment build (python -m objectpermanence_b200.build --exp skipl1 OPN_FUSED_SKIP_L1=1), which reads frames_boxes
from the tensors the first run left, checks that h2 is the same and times both."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
exp = ctypes.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), "libopnet_b200_skipl1.so"))
for name in ("opn_opnet_fwd", "opn_opnet_fwd_workspace_bytes"):
    getattr(exp, name).restype, getattr(exp, name).argtypes = _lib.SIGNATURES[name]
B, T, H1, H2 = 32, 300, 256, 512
f32 = dict(device=dev, dtype=torch.float32)
g = torch.Generator().manual_seed(3)
r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(dev)
boxes = torch.rand(B, T, 15, 6, generator=g).to(dev)
xproj1 = r(B, T, 4 * H1) * 0.5
w_hh1, w_pred, w_ih2, w_hh2 = r(4 * H1, H1) / H1 ** 0.5, r(15, H1) / H1 ** 0.5, r(4 * H2, 6) / H2 ** 0.5, r(4 * H2, H2) / H2 ** 0.5
hs1, g1, c1 = torch.empty(B, T, H1, **f32), torch.empty(B, T, 4 * H1, **f32), torch.empty(B, T, H1, **f32)
logits, probs, fb = torch.empty(B, 15, T, **f32), torch.empty(B, T, 15, **f32), torch.empty(B, T, 6, **f32)
g2, c2 = torch.empty(B, T, 4 * H2, **f32), torch.empty(B, T, H2, **f32)
hs2 = [torch.empty(B, T, H2, **f32) for _ in range(2)]
ws = torch.zeros(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
def call(which, out):
    rc = which.opn_opnet_fwd(B, T, H1, H2, boxes.data_ptr(), xproj1.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(),
                             w_hh2.data_ptr(), hs1.data_ptr(), g1.data_ptr(), c1.data_ptr(), logits.data_ptr(), probs.data_ptr(),
                             fb.data_ptr(), out.data_ptr(), g2.data_ptr(), c2.data_ptr(), ws.data_ptr(), ws.numel(), s)
    assert rc == 0, rc
def timed(which, out):
    for _ in range(2): call(which, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): call(which, out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
t_full = timed(lib, hs2[0])
t_skip = timed(exp, hs2[1])
print(f"shipped fused forward            : {t_full:.4f} ms = {t_full * 1e3 / T:.3f} us/frame")
print(f"without LSTM1 / head on these SMs: {t_skip:.4f} ms = {t_skip * 1e3 / T:.3f} us/frame   (h2 equal: max|d| = {(hs2[0] - hs2[1]).abs().max().item():.1e})")
