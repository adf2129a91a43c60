# usage: tools/sweep_variants.sh <samples> tag1 tag2 ...   (variants built by tools/build_variant.sh)
S=$1; shift
for v in "$@"; do
  LQGK_LIB_PATH=lqg_b200/csrc/variants/liblqgk_$v.so python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-secondary --samples $S > gpurun_out/sweep_$v.json 2> gpurun_out/sweep_$v.err
done
python - "$@" <<'PY'
import json, sys
for v in sys.argv[1:]:
    try:
        j=json.loads(open(f"gpurun_out/sweep_{v}.json").read().strip().splitlines()[-1])
        k=j["kernels"]
        print(v,"step %.2f  trial_fwd %.2f trial_rev %.2f"%(j["ms_per_step"],k["trial_fwd"]["ms_per_step"],k["trial_rev"]["ms_per_step"]))
    except Exception as e:
        print(v,"ERR",e, open(f"gpurun_out/sweep_{v}.err").read()[-400:])
PY
