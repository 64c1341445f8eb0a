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
# HLC PLINK / -gk 2 / covariates pins, test/dev_tests.rb:81-95 (kept as literals in tests/test_oracle_golden.py)
# LOCO / -nind pins: test/dev_tests.rb:57-77 (row 2 logl_H1, max p_wald), test/dev_test_suite.sh:121-153 (cXX / assoc shape oracles)
m = re.search(r'mouse_hs1940_loco\.assoc\.txt",\[\[2,9,"([0-9.e+-]+)"\],\s*\[:max,"p_wald","([0-9.e+-]+)"\]\]', rb)
sh = open(REF + "/test/dev_test_suite.sh").read()
loco = sh[sh.index("testCenteredRelatednessMatrixKLOCO1"):sh.index("testPlinkCenteredRelatednessMatrixKLOCO1")]
exp["mouse_loco"] = {"source": "test/dev_tests.rb:57-77; test/dev_test_suite.sh:121-153", "row2_logl_H1": m.group(1), "max_p_wald": m.group(2),
                     "cxx_lines": int(re.search(r'assertEquals "(\d+)" `wc -l < \$outfn`', loco).group(1)),
                     "assoc_lines": int(re.findall(r'assertEquals "(\d+)" `wc -l < \$outfn`', loco)[1]),
                     "cxx_head5": re.search(r'assertEquals "([0-9.]+)" `head -c 5', loco).group(1),
                     "cxx_sum2": re.search(r'assertEquals "([0-9.]+)" `perl', loco).group(1)}
# multivariate rows (-n 1 6 -lmm), demo.txt:62-66
mv_hdr = [i for i, ln in enumerate(demo) if ln.startswith("chr\trs") and "beta_1" in ln][0]
exp["mouse_mvlmm_rows"] = {"source": "example/demo.txt:62-66 (-n 1 6 -lmm): " + " ".join(demo[mv_hdr].split("\t")),
                           "rows": [demo[i].split("\t") for i in range(mv_hdr + 1, mv_hdr + 6)]}
json.dump(exp, open(__file__.replace("make_expected.py", "expected.json"), "w"), indent=1)
print(json.dumps(exp, indent=1)[:1500])
