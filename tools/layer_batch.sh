for b in 32 64 128 256; do
  timeout 300 python bench.py --no-cpu-baseline --lanes 1 --batch $b --steps 50 > gpurun_out/lb_$b.json 2> gpurun_out/lb_$b.err
  python - <<P
import json
d=json.load(open("gpurun_out/lb_$b.json"))
print($b, round(d["value"]), "trunk_ms", round(d["roofline_trunk"]["avg_ms"],4), [round(x*1000,1) for x in d["layer_ms"] if x])
P
done
