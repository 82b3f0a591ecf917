"""Turns the raw ncu outputs that gpurun brings back into the committed summaries.

    python profiles/summarize.py gpurun_out/launches_r01.csv gpurun_out/prof_dslash_r01.ncu-rep r01

writes profiles/launches_<tag>.txt (per-kernel share of the step, from the
`--metrics gpu__time_duration.sum --clock-control none` pass), profiles/dslash_ncu_<tag>.csv
(selected raw metrics of the `--set full` capture) and refreshes
profiles/dslash_ncu_summary.json (read by bench.py for roofline.traffic).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = os.path.join(HERE, "launches_%s.txt" % tag)
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s, %d launches, %.1f us total\n" % (os.path.basename(path), sum(v[0] for v in agg.values()), tot))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%6.2f%%  n=%5d  avg=%9.2f us  %s\n" % (100 * v[1] / tot, v[0], v[1] / v[0], k))
    print(open(out).read())


def full(path, tag):
    # path: an .ncu-rep, or the raw CSV page exported from it on the GPU box (ncu -i rep --page raw --csv)
    raw = open(path).read() if path.endswith(".csv") else \
        subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    out = os.path.join(HERE, "dslash_ncu_%s.csv" % tag)
    ti = hdr.index("gpu__time_duration.sum")
    # (launches of an already-stopped solver are ~8 us no-ops: not stencil measurements)
    rows = rows[:2] + [r for r in rows[2:] if len(r) == len(hdr) and float(r[ti].replace(",", "")) > 20.0]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(open(out).read())

    def val(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
    import re
    kernels = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        m = re.search(r"dslash_kernel<(double|float), (\d), (\d), (\d)>", name)
        mh = re.search(r"dslash_half_kernel<(\d), (\d), (\d)>", name)
        mm = re.search(r"dslash_mrhs_kernel<(double|float), (\d), (\d), (\d)>", name)
        if not m and not mh and not mm:
            continue
        tr = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        nrhs = 1
        if mh:
            prec, epi, mode, nc = 0, int(mh.group(1)), int(mh.group(2)), int(mh.group(3))
        elif mm:
            prec, epi, mode, nc = (2 if mm.group(1) == "double" else 1), int(mm.group(2)), 0, int(mm.group(4))
            nrhs = int(mm.group(3))
        else:
            prec, epi, mode, nc = (2 if m.group(1) == "double" else 1), int(m.group(2)), int(m.group(3)), int(m.group(4))
        if any(k["kernel"] == name for k in kernels):
            continue   # first launch of each variant
        kernels.append({"kernel": name, "prec": prec, "epilogue": epi, "nrhs": nrhs, "tag": tag,
                        "mode": mode, "long_reals": 2 * nc, "dram_bytes_per_launch": tr,
                        "dram_bytes_read": val(r, "dram__bytes_read.sum"),
                        "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                        "gpu_time_us": float(r[hdr.index("gpu__time_duration.sum")])})
    # kernels not in this capture keep the entry of the capture they were last seen in
    sp = os.path.join(HERE, "dslash_ncu_summary.json")
    if os.path.exists(sp):
        for k in json.load(open(sp)).get("kernels", []):
            if not any(n["kernel"] == k["kernel"] for n in kernels):
                kernels.append(k)
    js = {"tag": tag, "source": os.path.basename(path), "lattice": "32x32x32x64", "kernels": kernels}
    json.dump(js, open(sp, "w"), indent=1)
    print(json.dumps(js, indent=1))


if __name__ == "__main__":
    if os.path.exists(sys.argv[1]):
        launches(sys.argv[1], sys.argv[3])
    if os.path.exists(sys.argv[2]):
        full(sys.argv[2], sys.argv[3])
