"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"^void ", "", name)
    m = re.match(r"[\w:]+", name)
    base = m.group(0) if m else name
    if base.startswith("at::") and "elementwise" in base or base.startswith("at::native::"):
        inner = re.search(r"at::native::(\w+)", name)
        base = base + ("/" + inner.group(1) if inner else "")
    name = base[:70]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v_us = v / 1e3 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1e3)
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += v_us
total = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:70s} {n:8d} {us:10.1f} {us / total * 100:6.1f}% {us / n:8.1f}")
print(f"{'TOTAL':70s} {sum(v[0] for v in tot.values()):8d} {total:10.1f}")
