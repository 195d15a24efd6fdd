"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): the LAST full step found between two
sampler_uniform_kernel launches, per-kernel totals and shares."""
import csv, re, sys, collections
f = sys.argv[1]
lines = [l for l in open(f) if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [re.sub(r"^void ", "", r[ki]) for r in rows]
t = [float(r[vi]) / 1000.0 for r in rows]
# a step starts at the first pose_from_cam7_kernel... use weight_norm_kernel<0> (forward) as the step marker
marks = [i for i, n in enumerate(names) if "weight_norm_kernel<0>" in n and (i == 0 or "weight_norm_kernel<0>" not in names[i - 1])]
starts = [marks[0]] + [m for a, m in zip(marks, marks[1:]) if m - a > 50]
lo, hi = (starts[-2], starts[-1]) if len(starts) >= 2 else (0, len(rows))
if "--last" in sys.argv: lo, hi = starts[-1], len(rows)
agg = collections.OrderedDict()
for n, d in zip(names[lo:hi], t[lo:hi]):
    short = re.sub(r"\(.*$", "", n)
    short = re.sub(r"at::native::|native::|<unnamed>::", "", short)[:70]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += d
tot = sum(v[1] for v in agg.values())
print(f"# launches {hi - lo}, serialised {tot / 1000:.3f} ms  (rows {lo}..{hi})")
print("kernel,launches,ms,share_pct")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{v[0]},{v[1] / 1000:.4f},{100 * v[1] / tot:.2f}")
ours = sum(v[0] for k, v in agg.items() if k.startswith("nicer::"))
print(f"# nicer:: launches {ours}, torch/library launches {hi - lo - ours}")
