"""Regenerates tests/golden/expected.json from the reference's own documented outputs.
Run in the authoring container (needs /root/reference).  The values are the reference's
golden vectors for the -gk / -lmm path."""
import json
import re

REF = "/root/reference"
demo = open(REF + "/example/demo.txt").read().splitlines()
exp = {"source": {"mouse": "example/demo.txt:9-12,31-36,41-42", "bxd": "test/dev_tests.rb:42-43,53-54",
                  "counts": "test/performance/releases.org:11-16", "getab": "test/src/unittests-math.cpp:17-25"}}
# kinship 3x3 block, demo.txt:10-12
exp["mouse_K3"] = [[float(x) for x in demo[i].split()[:3]] for i in (9, 10, 11)]
# first five -lmm 1 rows, demo.txt:32-36
hdr = demo[30].split("\t")
rows = [dict(zip(hdr, demo[i].split("\t"))) for i in range(31, 36)]
exp["mouse_lmm1_rows"] = rows
exp["mouse_pve"] = float(re.search(r"= ([0-9.]+)", demo[40]).group(1))
exp["mouse_pve_se"] = float(re.search(r"= ([0-9.]+)", demo[41]).group(1))
exp["mouse_counts"] = {"ni_total": 1940, "ni_test": 1410, "ns_total": 12226, "ns_test": 10768}
rb = open(REF + "/test/dev_tests.rb").read()
exp["bxd_lmm2_row2_p_lrt"] = float(re.search(r'\[2,9,"([0-9.e+-]+)"\]', rb).group(1))
exp["bxd_max_p_lrt"] = float(re.search(r'\[:max,"p_lrt","([0-9.e+-]+)"\]', rb).group(1))
exp["bxd_lmm9_max_l_mle"] = float(re.search(r'\[:max,"l_mle","([0-9.e+-]+)"\]', rb).group(1))
exp["bxd_lmm2_assoc_words"] = 73180      # test/dev_test_suite.sh:83
ut = open(REF + "/test/src/unittests-math.cpp").read()
exp["getab"] = [[int(a), int(b), int(c), int(r)] for a, b, c, r in
                re.findall(r"GetabIndex\((\d+),\s*(\d+),\s*(\d+)\)\s*==\s*(\d+)", ut)]
json.dump(exp, open(__file__.replace("make_expected.py", "expected.json"), "w"), indent=1)
print(json.dumps(exp, indent=1)[:1500])
