import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["per_kernel"].items()}, 'loss',d["loss"], 'e2e', round(d["e2e"]["ms_per_step"],2))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-2000:])
