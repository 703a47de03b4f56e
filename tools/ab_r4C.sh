run() { tag=$1; b=$2; shift; shift; echo "== $tag batch=$b: $(env "$@" python bench.py --batch $b --steps 40 --warmup 5 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --vae-batch 0 --no-extra-legs 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"])')"; }
for b in 1 32; do
run oh1 $b WD_OUT_HEAD=1
run oh0 $b WD_OUT_HEAD=0
done
