#!/bin/bash
# Round-2 evidence run on one B200 (gpurun -- bash tools/evidence_r2.sh): the GPU test suite, smoke(),
# the default bench line, the ncu launch list of the same command and one `ncu --set full` capture of
# the fused gather.  Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand.
set -u
out=gpurun_out
mkdir -p $out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu -x > $out/r2_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $out/r2_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (default)"
timeout 600 python bench.py --steps 5 --warmup 3 > $out/r2_bench_n1.json 2> $out/r2_bench_n1.err; echo "rc=$?"
python - <<'EOF'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json"))
print(d["value"], d["unit"], d["ms_per_step"], {k: (round(v["ms"], 3), round(v.get("frac_of_peak", 0), 3)) for k, v in d["roofline"]["stages"].items()})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("frac_of_pcie"), "| roi_only", d["e2e_roi_only"]["value"], d["e2e_roi_only"].get("frac_of_pcie"))
print("parity", d["parity"]["ok"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("strong", json.dumps(d["strong_scaling"])[:600])
EOF
echo "== ncu launch list of the same command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"stitch_|roi_|ff_|chip_masks|bead_|bounding|mask_count" -c 80 --csv --log-file $out/r2_launches.csv \
  python bench.py --steps 5 --warmup 3 > $out/r2_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu --set full of the gather"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_gather_lists -s 1 -c 1 -f \
  -o $out/r2_gather_full python tools/gather_ncu.py 50 > $out/r2_ncu_gather.log 2>&1; echo "rc=$?"
ls -la $out | tail -12
