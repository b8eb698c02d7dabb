"""Where does the per-launch time of the fused forward go: host-side launch cost vs device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lstm_time.py")).read()
head, tail = src.split("# fused OPNet forward (LSTM1 + who-to-track + LSTM2 in one persistent kernel)")
exec(head.split("for B, H in itertools.product")[0] + "\nT = int(os.environ.get('TT', '300'))\n" + tail.split("fused(); torch.cuda.synchronize()")[0])
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); fused(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"T={T}: host call {1e3 * (t1 - t0):.3f} ms, until idle {1e3 * (t2 - t0):.3f} ms", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fused(); e1.record(); torch.cuda.synchronize()
print(f"T={T}: events around one call {e0.elapsed_time(e1):.3f} ms")
st = ws[:64].view(torch.int32).cpu().tolist()
print(f"sweeps per frame: h1 {st[8] / T:.2f}, h2 {st[9] / T:.2f}")
