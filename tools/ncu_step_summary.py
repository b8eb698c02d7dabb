"""Print the kernels of one bench step from an ncu launch list (gpu__time_duration.sum csv) and their shares."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
seq = []
for r in rows[start + 2:]:
    if len(r) <= vi:
        continue
    try:
        seq.append((r[ki], float(r[vi].replace(',', ''))))
    except ValueError:
        pass
short = lambda n: n.replace('void ', '').replace('opn::<unnamed>::', '').replace('(int)', '').replace('(bool)', '')
# one step = the launches between two consecutive fused forward kernels (cyclic: the x-projection contraction in
# front of the kernel is counted at the end); without the fused forward, between two LSTM1 forward recurrences
idx = [i for i, (n, v) in enumerate(seq) if 'opnet_fwd_fused' in n]
if len(idx) < 3:
    idx = [i for i, (n, v) in enumerate(seq) if 'lstm_fwd' in n and '256' in short(n)[:40]]
a, b = idx[1], idx[2]
tot = sum(v for _, v in seq[a:b])
for n, v in seq[a:b]:
    print(f"{v / 1000:9.1f} us {100 * v / tot:5.1f}%  {short(n)[:120]}")
print(f"step total {tot / 1e6:.3f} ms over {b - a} launches")
