# A/B of the group-sum pass 1 of the fused GroupNorm epilogue (DP_GN_P1=1 default / 0): kernel self-test with several work
# units per CTA pair (small grid), the whole -m gpu suite, then the loop on the same box with the switch on and off.
mkdir -p gpurun_out; cd diffpure_b200
echo "== selftest gn, 8 SMs"; DP_SELFTEST_SMS=8 timeout 120 ./selftest_gemm gn 2>&1 | grep -v "OK " | tail -6
cd ..
timeout 200 python -m pytest tests/test_gpu_parity.py -q -s -k "pair_tiles" 2>&1 | grep -E "pair tiles|passed|failed|assert" | head -6
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4
for p1 in 1 0; do DP_GN_P1=$p1 timeout 120 python bench.py --steps 3 --warmup 2 --no-secondary --no-gpu-eager --no-cpu-baseline --no-e2e > gpurun_out/ab_p1_$p1.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/ab_p1_$p1.log").read().strip().splitlines()[-1]); r=d["roofline"]
print("P1=$p1", round(d["value"],1), d["clocks"]["sm_mhz"], round(r["frac"],3), r["eval_ms_by_kind"]["gemm"], round(sum(r["eval_ms_by_kind"].values()),2))
PY
done
