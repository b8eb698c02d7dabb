"""smoke() of __graft_entry__ with a watchdog: the Python stack of every thread goes to stderr every WATCHDOG seconds
(default 60), so that a stall under a profiler shows where the host is waiting.  No rebuild (the library must be current)."""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
faulthandler.dump_traceback_later(float(os.environ.get("WATCHDOG", "60")), repeat=True)
t0 = time.time()
import __graft_entry__ as entry  # noqa: E402

print(f"[{time.time() - t0:6.1f}s] imported", flush=True)
import torch  # noqa: E402

print(f"[{time.time() - t0:6.1f}s] torch imported, cuda available: {torch.cuda.is_available()}", flush=True)
torch.zeros(4, device="cuda").add_(1).sum().item()
print(f"[{time.time() - t0:6.1f}s] first torch kernel done", flush=True)
entry.smoke()
print(f"[{time.time() - t0:6.1f}s] smoke done", flush=True)
