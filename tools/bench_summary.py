import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("ms/step %.2f  value %.2f  launches %d  e2e %s" % (d["ms_per_step"], d["value"], d["gpu_launches"], d.get("e2e",{}).get("value")))
tot=0
for k,v in sorted(d["roofline"]["breakdown"].items(), key=lambda kv:-kv[1]["ms_per_step"]):
    print("%-22s %4d %8.2f ms %8.1f" % (k, v["launches"], v["ms_per_step"], v.get("tflops") or 0)); tot+=v["ms_per_step"]
print("sum", tot)
