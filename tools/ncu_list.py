"""One line per captured launch of an `ncu --set full` report: duration, DRAM bytes, tensor-pipe activity, hit rates.
    python tools/ncu_list.py gpurun_out/r02_block_c256_bf16.ncu-rep "header text" > profiles/...txt"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
if len(sys.argv) > 2:
    print(sys.argv[2])
for r in rows[2:]:
    g = lambda n: float(r[ix[n]].replace(",", ""))
    u = lambda n: units[ix[n]]
    mb = lambda n: g(n) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u(n)]
    us = g("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u("gpu__time_duration.sum"), 1)
    print("{:30s} {:7.1f} us  DRAM {:6.1f}+{:6.1f} MB ({:4.1f} % of peak)  tensor pipe {:4.1f} %  L1 hit {:4.1f} %  L2 hit {:4.1f} %  issue {:4.1f} %".format(
        r[ix["Kernel Name"]].split("(")[0].split("::")[-1][:30], us, mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum"),
        g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        g("l1tex__t_sector_hit_rate.pct"), g("lts__t_sector_hit_rate.pct"), g("smsp__issue_active.avg.pct_of_peak_sustained_active")))
