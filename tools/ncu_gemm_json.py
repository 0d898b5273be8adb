"""Summarise an `ncu --set full` capture of the four forward GEMMs of an encoder block as the JSON `bench.py` reads
`roofline.traffic` from:   python tools/ncu_gemm_json.py gpurun_out/prof_gemm_r02_paired.ncu-rep profiles/out.json "<source note>"
(reads the report with `ncu -i <rep> --page raw --csv`)."""
import csv
import io
import json
import subprocess
import sys

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in reversed(list(enumerate(hdr)))}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "%": 1.0}


def val(r, name, to=1.0):
    i = col[name]
    return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0) / to


# launch order inside a block's forward: qkv, proj, fc1, fc2 (-s 30 lands on fc1 of a block); identified by DRAM pattern
NAMES = {(4608, 1152): ("fc1+GELU fwd (M=32768,N=4608,K=1152)", 690e6),
         (1152, 4608): ("fc2+gate+residual fwd (M=32768,N=1152,K=4608)", 690e6),
         (3456, 1152): ("qkv fwd (M=32768,N=3456,K=1152)", 310e6),
         (1152, 1152): ("proj+gate+residual fwd (M=32768,N=1152,K=1152)", 456e6)}
ORDER = [(4608, 1152), (1152, 4608), (3456, 1152), (1152, 1152)]
launches = []
for r, key in zip(data, ORDER):
    name, alg = NAMES[key]
    rd, wr = val(r, "dram__bytes_read.sum", 1e6), val(r, "dram__bytes_write.sum", 1e6)
    launches.append({
        "gemm": name,
        "kernel": r[col["Kernel Name"]].replace("void ", "").split("(")[0].replace(" ", ""),
        "duration_us": val(r, "gpu__time_duration.sum"),
        "dram_read_MB": rd, "dram_write_MB": wr, "dram_bytes": (rd + wr) * 1e6, "algorithmic_bytes": alg,
        "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_memory_throughput_pct": val(r, "sm__memory_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_to_sm_read_GB": val(r, "l1tex__m_xbar2l1tex_read_bytes.sum", 1e9),
        "lts_throughput_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
        "grid": val(r, "launch__grid_size"),
    })
json.dump({"source": note, "launches": launches}, open(out, "w"), indent=1)
for l in launches:
    print(f"{l['gemm']:52s} {l['duration_us']:7.1f} us  DRAM {l['dram_bytes'] / 1e6:7.1f} MB (alg {l['algorithmic_bytes'] / 1e6:.0f})"
          f"  tensor {l['tensor_pipe_active_pct']:.1f} %  smem/L1 {l['sm_memory_throughput_pct']:.1f} %")
