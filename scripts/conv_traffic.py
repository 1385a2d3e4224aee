"""profiles/rNN_conv_tc_traffic.json from an ncu CSV (dram__bytes_read/write + duration per conv_tc launch of one
denoising step) and the event-timed per-launch shape dump (pd_prof_dump) of the same step.
    python scripts/conv_traffic.py gpurun_out/conv_tc_dram.csv gpurun_out/gemm_shapes.csv profiles/r02_conv_tc_traffic.json"""
import csv, json, sys
ncu, shapes, out = sys.argv[1:4]
rows = list(csv.DictReader(l for l in open(ncu) if not l.startswith("==")))
rd = wr = t = 0.0
ids = set()
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "dram__bytes_read.sum":
        rd += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]; ids.add(r["ID"])
    elif r["Metric Name"] == "dram__bytes_write.sum":
        wr += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    elif r["Metric Name"] == "gpu__time_duration.sum":
        t += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}[u]
alg = 0.0
n = 0
for r in csv.DictReader(open(shapes)):
    M, N, K, ks = int(r["M"]), int(r["N"]), int(r["K"]), int(r["ksize"])
    alg += (M * K / (ks * ks) + M * N + N * K) * 2
    n += 1
json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc over one "
                     "eager denoising step (scripts/gpu_round2_profile.sh), config 2 (B_eff 16, 64x64 latent)",
           "launches_per_denoise_step": len(ids), "dram_bytes_read_per_denoise_step": rd, "dram_bytes_write_per_denoise_step": wr,
           "dram_bytes_per_denoise_step": rd + wr, "algorithmic_bytes_per_denoise_step": alg, "launches_in_shape_dump": n,
           "kernel_time_ms_under_ncu": t,
           "note": "algorithmic = sum over launches of (M*K/taps + M*N + N*K)*2 B (A read once per tap set, output, weights)"},
          open(out, "w"), indent=1)
print(open(out).read())
