"""Extract the headline metrics of every launch in an .ncu-rep (ncu --set full) as JSON lines."""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "dur_us", "dram__bytes_read.sum": "dram_rd_MB", "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "lts__t_sectors_srcunit_tex.sum": "l2_to_l1_sectors", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid", "launch__block_size": "block", "smsp__inst_executed.sum": "warp_insts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_bytes.sum": "l2_bytes_MB",
}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")][:60]}
        for m, k in WANT.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if k.endswith("_MB"):
                    v = v / 1e6 if u in ("byte", "bytes") else (v / 1e3 if u == "Kbyte" else (v if u == "Mbyte" else v * 1e3))
                if k == "dur_us":
                    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
                rec[k] = round(v, 3)
        print(json.dumps(rec))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("#", p)
        main(p)
