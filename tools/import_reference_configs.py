#!/usr/bin/env python3
"""Re-emit the reference's etc/ configs (YAML + SRDF) into inria_wbc_b200/etc/.

The batched path must consume the reference's *input format* unchanged
(SURVEY.md §2.1 row 5), but /root/reference does not exist on the GPU box, so
the configs the five BASELINE configs need are normalised here (PyYAML load ->
dump, insertion order kept because task order = Hessian summation order,
/root/reference/src/controllers/pos_tracker.cpp:167-182) and committed.
Only data is converted; no reference source code is copied.

usage: python tools/import_reference_configs.py [/root/reference]
"""
import os
import re
import sys
import yaml

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "inria_wbc_b200", "etc")

YAMLS = {
    "franka": ["pos_tracker.yaml", "tasks.yaml", "cartesian_line.yaml"],
    "talos": ["talos_pos_tracker.yaml", "tasks.yaml", "frames.yaml", "squat.yaml", "arm.yaml",
              "walk.yaml", "walk_on_spot.yaml", "move_com.yaml", "move_feet.yaml", "clapping.yaml", "align.yaml",
              "collisions/talos_collisions_margin.yaml", "collisions/talos_collisions.yaml",
              "stab_double_support.yaml", "stab_single_support.yaml", "stab_fixed_base.yaml", "collision_thresholds.yaml",
              "talos_pos_tracker_real_robot.yaml"],
    "icub": ["humanoid_pos_tracker.yaml", "tasks.yaml", "frames.yaml", "squat.yaml", "arm.yaml", "walk.yaml", "walk_on_spot.yaml", "clapping.yaml", "align.yaml",
             "stab_double_support.yaml", "stab_single_support.yaml", "stab_fixed_base.yaml", "collision_thresholds.yaml"],
    "tiago": ["pos_tracker.yaml", "tasks.yaml", "frames.yaml", "cartesian_line.yaml", "collisions.yaml"],
}
SRDFS = {
    "franka": ["configurations.srdf"],
    "talos": ["configurations.srdf"],
    "icub": ["configurations.srdf"],
    "tiago": ["configurations-elbow-down.srdf", "configurations-elbow-up.srdf"],
}


class Str(str):
    pass


def _mask_safe_load(text):
    # masks such as `mask: 000110` must stay strings (yaml-cpp reads them with as<std::string>())
    text = re.sub(r"(mask\s*:\s*)([01]+)\s*(#.*)?$", r'\1"\2"', text, flags=re.M)
    # etc/franka/pos_tracker.yaml ends with an unterminated `collision_path: "` (the key is unused
    # because check_model_collisions is false there): close the quote
    text = re.sub(r'(collision_path\s*:\s*)"\s*\Z', r'\1""', text)
    return yaml.safe_load(text)


def main():
    for robot, files in YAMLS.items():
        os.makedirs(os.path.join(OUT, robot), exist_ok=True)
        for f in files:
            src = os.path.join(REF, "etc", robot, f)
            if not os.path.exists(src):
                print("skip (absent):", src)
                continue
            node = _mask_safe_load(open(src).read())
            if isinstance(node, dict) and "CONTROLLER" in node:
                node["CONTROLLER"]["base_path"] = "@ETC@/" + robot   # resolved by the loader
            os.makedirs(os.path.dirname(os.path.join(OUT, robot, f)), exist_ok=True)
            with open(os.path.join(OUT, robot, f), "w") as fo:
                fo.write("# normalised from the reference's etc/%s/%s by tools/import_reference_configs.py\n" % (robot, f))
                yaml.safe_dump(node, fo, sort_keys=False, default_flow_style=None, width=100)
        for f in SRDFS[robot]:
            src = os.path.join(REF, "etc", robot, f)
            txt = open(src).read()
            groups = re.findall(r'<group_state\s+name="([^"]+)"[^>]*>(.*?)</group_state>', txt, flags=re.S)
            with open(os.path.join(OUT, robot, f), "w") as fo:
                fo.write('<?xml version="1.0"?>\n<!-- normalised from the reference\'s etc/%s/%s -->\n<robot name="%s">\n' % (robot, f, robot))
                for gname, body in groups:
                    fo.write('  <group_state name="%s" group="all">\n' % gname)
                    for jn, val in re.findall(r'<joint\s+name="([^"]+)"\s+value="([^"]*)"', body):
                        fo.write('    <joint name="%s" value="%s"/>\n' % (jn, " ".join(val.split())))
                    fo.write("  </group_state>\n")
                fo.write("</robot>\n")
    print("configs written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
