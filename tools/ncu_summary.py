"""Summarise an .ncu-rep (read here, no GPU): key raw metrics per kernel + top stall instructions (source page)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print("=== ", d[ix["Kernel Name"]][:110])
        for w in WANT:
            if w in ix:
                print(f"  {w:72s} {d[ix[w]]:>18s} {units[ix[w]]}")
    return hdr


def source(rep, kid, top=28):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix["stall_wait"]]
    seen, uniq = set(), []
    for r in data:
        if r[ix["Address"]] in seen:
            continue
        seen.add(r[ix["Address"]])
        uniq.append(r)
    iv = lambda x: int(x) if x.isdigit() else 0  # noqa: E731
    tot = sum(iv(r[ix["# Samples"]]) for r in uniq)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
    agg = {s: sum(iv(r[ix[s]]) for r in uniq) for s in stalls}
    print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(uniq, key=lambda r: -iv(r[ix["# Samples"]]))[:top]:
        st = sorted(((iv(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"  {r[ix['# Samples']]:>6s} {r[ix['Source']][:86]:86s} {st}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    for kid in sys.argv[2:]:
        source(rep, kid)
