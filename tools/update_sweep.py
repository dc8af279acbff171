"""BASELINE config 5: fused superposition-update kernel alone, B = 1..512 at 256^2 (D = 65536), M = 2.
MODE=rotating (default; back-to-back launches on rotating buffer sets > 4 x L2) or MODE=flush (L2 flush per launch)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
hbm = bench.peaks()[0]
rot = os.environ.get("MODE", "rotating") == "rotating"
Bs = [int(b) for b in os.environ.get("BS", "1,2,4,8,16,32,64,128,256,512").split(",")]
rows = []
print("mode", "rotating" if rot else "flush", "steps", os.environ.get("SDD_UPD_STEPS", "default"))
for B in Bs:
    it = 200 if rot else 20
    gb, ms = bench.update_roofline(S, dev, B, 65536, iters=it, rotating=rot)
    gb2, ms2 = bench.update_roofline(S, dev, B, 65536, iters=it, noise=True, rotating=rot)
    rows.append({"B": B, "mode": "rotating" if rot else "flush", "philox_us": ms * 1000, "philox_gbs": gb, "philox_frac": gb / hbm, "noise_us": ms2 * 1000, "noise_gbs": gb2, "noise_frac": gb2 / hbm})
    print(f"update B={B:4d}: philox(16 B/el) {ms*1000:7.1f} us {gb:7.1f} GB/s ({gb/hbm:.3f}) | noise tensor(20 B/el) {ms2*1000:7.1f} us {gb2:7.1f} GB/s ({gb2/hbm:.3f})", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(os.environ.get("OUT", "gpurun_out/update_sweep.json"), "w"), indent=1)
