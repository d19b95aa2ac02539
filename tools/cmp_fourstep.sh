for n in 8192 16384 32768 65536 131072 262144; do for sz in 26 30; do for f in CS16 CF32; do
c=X:$f:$n:1:$sz
a=$(python tools/sweep.py $c 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('%.1f' % (d['msamples_s']/1e3))")
b=$(SP_FOURSTEP=hbm python tools/sweep.py $c 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('%.1f' % (d['msamples_s']/1e3))")
echo "$c big $a GS/s  hbm-scratch $b GS/s"
done; done; done
