"""C4 sub-record alone (ViT-L/14 @224, 8 patients per step): value and the attention kernel's time inside the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
D = bench.Dist()
r = bench.bench_extraction(D, "C4", 10, 3, profile=True, cpu_slices=2, patients_per_step=8)
a = [v for k, v in r["roofline"]["per_kernel"].items() if k.startswith("attn")][0]
print(f"C4 {r['value']:.0f} slices/s, {r['ms_per_step']:.2f} ms/step, attention {a['ms_per_launch']:.4f} ms/launch", flush=True)
