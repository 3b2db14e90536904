import json,sys
d=json.load(open(sys.argv[1]))
print("frames/s %.0f  ms/step %.2f  K2 frac %.4f  K2 ms %.3f  parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d.get("parity_check")))
