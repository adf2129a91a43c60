for ck in 1 2 4; do
  LQGK_LIB_PATH=lqg_b200/csrc/variants/liblqgk_ck$ck.so python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-secondary --samples 16384 > gpurun_out/r02b_ck$ck.json 2> gpurun_out/r02b_ck$ck.err
done
python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-secondary --samples 16384 > gpurun_out/r02b_ck8.json 2> gpurun_out/r02b_ck8.err
python - <<'PY'
import json
for ck in (1,2,4,8):
    try:
        j=json.loads(open(f"gpurun_out/r02b_ck{ck}.json").read().strip().splitlines()[-1])
        k=j["kernels"]
        print("CK",ck,"step %.2f  trial_fwd %.2f trial_rev %.2f"%(j["ms_per_step"],k["trial_fwd"]["ms_per_step"],k["trial_rev"]["ms_per_step"]))
    except Exception as e:
        print("CK",ck,"ERR",e, open(f"gpurun_out/r02b_ck{ck}.err").read()[-400:])
PY
