#!/usr/bin/env python3
"""Extract the reference's own known-answer values for the hot path into tests/golden/.

Source: /root/reference/tests/ref_test_franka.yaml (written by the reference's test binary with
--generate_ref, tests/test_franka.cpp:243-250,434-437).  /root/reference does not exist on the GPU
box, so the numbers are committed as a small JSON fixture together with this script.
"""
import json
import os
import sys

import yaml

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_test_franka.json")

doc = yaml.safe_load(open(os.path.join(REF, "tests", "ref_test_franka.yaml")))
node = doc["../../etc/franka/pos_tracker.yaml"]["../../etc/franka/cartesian_line.yaml"]
out = {"source": "tests/ref_test_franka.yaml", "timestamp": str(doc["timestamp"]), "tolerance": 5e-3,
       "duration_s": 10, "dt": 0.001, "cases": {}}
for act, v in node.items():
    leaf = v["fcl"]["franka/franka.urdf"]
    out["cases"][act] = {"ee_error_tsid": leaf["ee"][0], "ee_error_dart": leaf["ee"][1], "time": leaf["time"]}
json.dump(out, open(OUT, "w"), indent=1)
print(json.dumps(out, indent=1))
