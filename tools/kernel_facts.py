#!/usr/bin/env python3
"""profiles/kernel_facts.json from `ncu --set full` captures of the dominant kernel of each BASELINE config.

    python tools/kernel_facts.py <git sha of the captured build> c1=gpurun_out/r2f_c1.ncu-rep c2=... c5=gpurun_out/r2f_c5n8.ncu-rep

bench.py reads the file for what only a profiler sees -- DRAM bytes per launch (`roofline.traffic`) and active lanes per warp
instruction -- and names the capture it came from.  Also prints the one-line summaries kept in profiles/."""
import csv, io, json, pathlib, subprocess, sys

sha = sys.argv[1]
out = {}
for arg in sys.argv[2:]:
    key, rep = arg.split("=")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))

    def num(name, scale=None):
        v = float(d[name].replace(",", ""))
        unit = u.get(name, "")
        if scale == "bytes":
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if scale == "ms":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(unit, 1)
        return v

    rd, wr = num("dram__bytes_read.sum", "bytes"), num("dram__bytes_write.sum", "bytes")
    out[key] = {
        "kernel": d["Kernel Name"].split("(")[0].replace("void ", "").replace("rtcu_dev::", ""),
        "dram_bytes_per_launch": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
        "active_lanes": round(num("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
        "issue_active_pct": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
        "fma_pipe_cycles_pct": round(num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), 1),
        "warp_instructions": int(num("smsp__inst_executed.sum")),
        "ms_under_ncu": round(num("gpu__time_duration.sum", "ms"), 3),
        "registers": int(num("launch__registers_per_thread")),
        "l1_hit_pct": round(num("l1tex__t_sector_hit_rate.pct"), 1), "l2_hit_pct": round(num("lts__t_sector_hit_rate.pct"), 1),
        "source": f"ncu --set full --clock-control none, one launch, library built at git {sha}: profiles/{pathlib.Path(rep).stem}_summary (kernel_facts.json)",
    }
    print(key, json.dumps(out[key]))
path = pathlib.Path(__file__).resolve().parent.parent / "profiles" / "kernel_facts.json"
path.write_text(json.dumps(out, indent=1) + "\n")
print("wrote", path)
