import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d["value"], d["roofline"]["frac"], d["roofline"]["deliver_ms_total"], d["roofline"]["update_ms_total"])
