#!/bin/bash
# copies the outputs of tools/gpu_final1.sh / gpu_final2.sh (gpurun_out/f1_*, f2_*) into profiles/ (run in the build container)
set -e
cd "$(dirname "$0")/.."
HDR="round 2 final code (sw128 weight gradients, 20-SM depth lane, 512-thread forward program with three epilogue groups): RADAR_DEPTH_B200_GRAPHS=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 800 -c 520 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing"
cp gpurun_out/f2_launches.csv profiles/r02_launches.csv
python tools/ncu_launch_summary.py profiles/r02_launches.csv "$HDR" --traffic-json profiles/r02_conv_traffic.json > profiles/r02_launches_summary.txt
python - <<'PY' >> profiles/r02_launches_summary.txt
import sys
sys.path.insert(0,'tools')
import ncu_launch_summary as n
L=n.parse('profiles/r02_launches.csv')
tot=sum(l['rd']+l['wr'] for l in L)
print(f"\n# DRAM bytes over the {len(L)} profiled launches (~2 training steps): {tot/1e9:.2f} GB = {tot/2e9:.2f} GB per step (round 1 / start of round 2: 12.5 GB per step)")
PY
python tools/ncu_summary.py gpurun_out/f2_full.ncu-rep "r02 final" > /tmp/ncu_full.txt 2>&1
python - <<'PY'
order=[l.split()[1] for l in open('gpurun_out/f2_full.log') if l.startswith('replayed')]
txt=open('/tmp/ncu_full.txt').read()
blocks=txt.split("== r02 final: ")
assert len(blocks)-1==len(order),(len(blocks),len(order))
out=["# ncu --set full --clock-control none --import-source on --profile-from-start off python tools/run_launch.py 1 16 <names>  (round 2 final code;",
     "# one eager launch of each named launch of the b=16 latefusion step after a full warm-up step; tools/ncu_summary.py)",
     "# launches inside the encoder run with their lane's SM budget (grid 128 = RGB chain, 20 = depth chain)",""]
for name,b in zip(order,blocks[1:]):
    out.append(f"== {name}: "+b.rstrip("\n"))
open('profiles/r02_ncu_full.txt','w').write("\n".join(out)+"\n")
PY
cp gpurun_out/f1_bench_n1.json profiles/r02_bench_n1_final.json
cp gpurun_out/f1_bench_multistage.json profiles/r02_bench_multistage_b8_final.json
cp gpurun_out/f1_bench_fp32.json profiles/r02_bench_fp32_parity_mode_final.json
cp gpurun_out/f1_bench_reference.json profiles/r02_bench_reference_arm_final.json
cp gpurun_out/f1_per_launch_latefusion.txt profiles/r02_per_launch_latefusion_final.txt
cp gpurun_out/f1_per_launch_multistage.txt profiles/r02_per_launch_multistage_final.txt
cp gpurun_out/f1_eval_latency.txt profiles/r02_eval_latency_final.txt
cp gpurun_out/f1_wgrad_sw128_check.txt profiles/r02_wgrad_sw128_check.txt
python tools/sass_summary.py > /dev/null
tail -4 profiles/r02_launches_summary.txt
