import json,sys
d=json.loads(sys.stdin.read()); e=d["e2e"]
print(sys.argv[1], round(d["value"]), round(e["value"]), round(e["batch"]["value"]))
for k,v in d["roofline"]["per_kernel"].items(): print("   %-34s %7.1f us x %.2f/frame" % (k, v["avg_us"], v["launches_per_frame"]))
print("   one frame per launch:")
for k,v in d["roofline"]["per_kernel_one_frame_per_launch"].items(): print("   %-34s %7.1f us x %.2f/frame" % (k, v["avg_us"], v["launches_per_frame"]))
