PREV=$PWD/scripts/probe/prev_libcoevonet_b200.so
for i in 1 2; do
echo -n "new : "; timeout 200 python scripts/time_roles.py 2>/dev/null | head -1 | cut -d: -f2 | cut -d, -f1
echo -n "prev: "; COEVONET_LIB=$PREV timeout 200 python scripts/time_roles.py 2>/dev/null | head -1 | cut -d: -f2 | cut -d, -f1
done
run() { echo -n "== $1: "; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'])"; }
run new CEV_X=1
run prev COEVONET_LIB=$PREV
run new CEV_X=1
run prev COEVONET_LIB=$PREV
