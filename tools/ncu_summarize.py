"""Summaries of ncu reports / launch lists for profiles/ (run in the dev container on files from gpurun_out/).
  python tools/ncu_summarize.py launches <launch list csv>
  python tools/ncu_summarize.py full <report.ncu-rep> [kernel regex]"""
import collections, csv, io, re, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic",
           # shared-memory side (round 2): tensor-core operand reads (1 wavefront = 128 B, 1 per clock per SM at peak), LSU traffic, uniform pipe
           "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            try:
                agg.setdefault(re.sub(r"^void ", "", r[ki])[:50], []).append(float(r[vi].replace(",", "")) / 1e3)
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print("%-50s n=%4d mean %8.1f us  min %7.1f max %8.1f  share %5.1f%%" % (k, len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))


def full(path, rx=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    seen = collections.Counter()
    for r in rows[2:]:
        name = r[ki]
        if rx and not re.search(rx, name):
            continue
        seen[name] += 1
        if seen[name] > 2:
            continue
        print("--- %s  (launch id %s)" % (name[:100], r[0]))
        for m in METRICS:
            if m in h:
                i = h.index(m)
                print("  %-70s %-16s %s" % (m, units[i], r[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
