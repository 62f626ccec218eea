cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,serial,power.limit,temperature.gpu --format=csv,noheader
for rep in 1 2 3; do
echo "lanes=1"; python bench.py --steps 40 --warmup 5 --lanes 1 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py | cut -c1-60
echo "lanes=2 tune"; python bench.py --steps 40 --warmup 5 --lanes 2 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py | cut -c1-60
echo "lanes=2 env"; RRV_NO_PDL=1 python bench.py --steps 40 --warmup 5 --lanes 2 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py | cut -c1-60
done
