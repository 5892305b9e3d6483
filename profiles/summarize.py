#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv   > profiles/r1_launches.txt
  python profiles/summarize.py kernel   gpurun_out/realize_r3.ncu-rep > profiles/r1_realize_v3.txt
  python profiles/summarize.py metrics  gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/ncu_metrics.json
        (per-kernel DRAM bytes / instruction counts per launch that bench.py quotes as `roofline.traffic`)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [ii for ii, rr in enumerate(rows) if rr and rr[0] == "ID"][0]
    head = rows[hdr]
    ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = collections.OrderedDict()
    for rr in rows[hdr + 1:]:
        if len(rr) <= vi:
            continue
        val = float(rr[vi].replace(",", ""))
        val = val / 1e6 if rr[ui] == "ns" else (val / 1e3 if rr[ui].startswith("us") else val)
        ent = agg.setdefault(rr[ki][:90], [0, 0.0])
        ent[0] += 1
        ent[1] += val
    tot = sum(vv[1] for vv in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source {path}")
    print(f"# total device time of all launches: {tot:.3f} ms (cold-cache, serialised: compare SHARES)")
    print(f"{'ms':>10s} {'calls':>6s} {'share':>7s}  kernel")
    for kk, (nn, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{tt:10.3f} {nn:6d} {100*tt/tot:6.1f}%  {kk}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none ; source {path}")
    for rr in rows[2:]:
        name = rr[head.index("Kernel Name")]
        print(f"\n## {name}")
        for hh, uu, vv in zip(head, units, rr):
            if hh in KEYS:
                print(f"{hh:85s} {vv:>22s} {uu}")


def metrics(paths):
    import json
    short = {"realize_kernel": "loudest_draw", "realize_quad_kernel": "loudest_draw_params", "dbn_2pwl_kernel": "dbn_2pwl",
             "bin_kernel": "integrate_strain", "norm_2pwl_kernel": "norm_2pwl", "density_kernel": "density",
             "ct_gradients_kernel": "scatter_gradients", "ct_eval_kernel": "scatter_ct_eval"}
    res = {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        head = rows[0]

        def val(rr, key):
            return float(rr[head.index(key)].replace(",", "")) if key in head else None
        for rr in rows[2:]:
            name = rr[head.index("Kernel Name")]
            key = next((vv for kk, vv in sorted(short.items(), key=lambda kv: -len(kv[0])) if kk in name), None)
            if key is None:
                continue
            units = rows[1]

            def to_bytes(metric):
                vv, uu = val(rr, metric), units[head.index(metric)]
                return vv * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[uu]
            res[key] = {
                "kernel": name.split("(")[0], "source": path.replace("gpurun_out/", ""),
                "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                "warp_inst_per_launch": val(rr, "smsp__inst_executed.sum"),
                "issue_active_pct": val(rr, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "ms_under_ncu": val(rr, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(
                    units[head.index("gpu__time_duration.sum")], 1.0),
            }
    # the hash of the draw kernel's sources at capture time: bench.py flags a capture of an older kernel version
    import hashlib
    import pathlib
    hh = hashlib.sha1()
    for name in ("holo_realize.cu", "holo_rng.cuh"):
        hh.update((pathlib.Path(__file__).resolve().parents[1] / "holodeck_b200" / "csrc" / name).read_bytes())
    for ent in res.values():
        ent["source_hash"] = hh.hexdigest()[:12]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "metrics":
        metrics(sys.argv[2:])
    else:
        {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
