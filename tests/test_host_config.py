"""Host logic: config readers, stack compiler, C-ABI surface (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import inria_wbc_b200 as iw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "iwbc_batch.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(iwbc_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    L = C.CDLL(iw.lib_path())
    for n in sorted(names):
        assert hasattr(L, n), "libiwbc_b200.so does not export " + n
    assert iw.lib().iwbc_abi_version() == 1


def test_qp_shapes_match_survey(franka, talos, tiago):
    # Franka n=9 (15 rows); Talos-full DS n=74 me=18 mi=244 rows=97; SS n=62 me=12 mi=210 rows=91
    assert (franka.nq, franka.nv, franka.na) == (9, 9, 9)
    assert franka.qp_shape() == (9, 0, 0, 15)
    assert (talos.nq, talos.nv, talos.na, talos.ncontacts) == (51, 50, 44, 2)
    assert talos.qp_shape() == (74, 18, 244, 97)
    assert talos.qp_shape(0b01) == (62, 12, 210, 91)
    assert talos.qp_shape(0) == (50, 6, 176, 85)
    assert tiago.qp_shape() == (12, 0, 24, 25)


def test_joint_order_is_pinocchio_like(talos):
    dofs = talos.names("dofs")
    assert dofs[:6] == ["leg_left_%d_joint" % i for i in range(1, 7)]
    assert dofs[12:14] == ["torso_1_joint", "torso_2_joint"]
    assert dofs[14] == "arm_left_1_joint" and dofs[-2:] == ["head_1_joint", "head_2_joint"]
    # the order of etc/talos/configurations.srdf (which is pinocchio's): gripper sub-tree, depth first, alphabetical
    assert dofs[21:28] == ["gripper_left_inner_double_joint", "gripper_left_fingertip_1_joint", "gripper_left_fingertip_2_joint",
                           "gripper_left_inner_single_joint", "gripper_left_fingertip_3_joint", "gripper_left_joint",
                           "gripper_left_motor_single_joint"]
    assert len(talos.names("mimics")) == 12
    assert talos.names("tasks")[:3] == ["head", "head_pitch", "head_yaw"]
    assert talos.names("contacts") == ["contact_lfoot", "contact_rfoot"]


def test_reference_slots_and_defaults(talos, franka):
    r = talos.default_refs()
    assert r.shape == (talos.nref,) and talos.nref == 8 * 24 + 9 + 3 * 44 + 18 + 2 * 32  # a contact block: SE3 sample (24) + f_ref (6) + weight-vector index + pad
    off, n = talos.ref_slot("com")
    assert n == 9
    lf, _ = talos.ref_slot("lf")
    rf, _ = talos.ref_slot("rf")
    # init_com: CoM reference moved to the middle of the feet (humanoid_pos_tracker.cpp:56-62)
    np.testing.assert_allclose(r[off:off + 2], 0.5 * (r[lf:lf + 2] + r[rf:rf + 2]), atol=1e-15)
    # contact references = ankle joint placements = the lf/rf task references
    cl, _ = talos.ref_slot("contact_lfoot")
    np.testing.assert_array_equal(r[cl:cl + 24], r[lf:lf + 24])
    po, pn = talos.ref_slot("posture")
    assert pn == 3 * 44
    np.testing.assert_allclose(r[po:po + 44], talos.q0()[7:], atol=0)
    # rotation part of an SE3 sample is column-major and orthonormal
    eo, _ = franka.ref_slot("ee")
    R = franka.default_refs()[eo + 3:eo + 12].reshape(3, 3).T
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    with pytest.raises(iw.IwbcError):
        talos.ref_slot("bounds")


def test_quaternion_is_normalised(talos):
    q0 = talos.q0()
    assert abs(np.linalg.norm(q0[3:7]) - 1) < 1e-15


def test_talos_hip_roll_bounds_are_tightened(talos):
    S = talos.struct()
    dofs = talos.names("dofs")
    q0 = talos.q0()[7:]
    roll = float(np.float32(14.32) / np.float32(180)) * np.pi
    for n in ("leg_left_2_joint", "leg_right_2_joint"):
        i = dofs.index(n)
        assert S.q_lo[i] == q0[i] - roll and S.q_hi[i] == q0[i] + roll


def test_gains_and_levels(talos):
    S = talos.struct()
    names = talos.names("tasks")
    t = S.tasks[names.index("lh")]
    assert t.type == 0 and t.level == 1 and t.nrows == 6 and t.weight == 10.0
    assert t.kp[0] == 30.0 and t.kd[0] == 2 * np.sqrt(30.0)
    t = S.tasks[names.index("torso")]
    assert list(t.mask[:6]) == [0, 0, 0, 1, 1, 0] and t.nrows == 2
    assert S.tasks[names.index("bounds")].level == 0 and S.tasks[names.index("actuation_bounds")].level == 0
    sc = S.tasks[names.index("self_collision-left")]
    assert sc.nav == 12 and sc.kp[0] == 50.0 and sc.kd[0] == 250.0
    k = -np.log((1 - 1e-5) ** (-1 / 0.2) - 1) / 0.02
    assert abs(sc.sc_k - k) < 1e-9 * k
    c = S.contacts[0]
    assert c.ub[16] == 1500.0 and c.lb[16] == 5.0 and c.lb[0] == -1e10 and c.w_forcereg == 1e-3
    # level-0 order: base dynamics, bounds, actuation bounds, then per contact motion + force
    kinds = [(S.l0[i].kind, S.l0[i].index) for i in range(S.nl0)]
    assert kinds[0][0] == 0 and [k for k, _ in kinds[1:]] == [4, 4, 1, 2, 1, 2]


def test_overrides_and_errors(tmp_path):
    P = iw.Problem("franka/pos_tracker.yaml", overrides={"dt": 0.002})
    assert P.struct().dt == 0.002
    with pytest.raises(iw.IwbcError, match="ref_map"):
        iw.Problem("franka/pos_tracker.yaml", overrides={"ref_config": "nope"})
    with pytest.raises(iw.IwbcError, match="eiquadprog"):
        iw.Problem("franka/pos_tracker.yaml", overrides={"solver": "qpmad"})
    with pytest.raises(iw.IwbcError):
        iw.Problem("franka/does_not_exist.yaml")
    # a task file with an unknown frame / a missing key fails with the key in the message (IWBC_CHECK behaviour)
    bad = tmp_path / "tasks.yaml"
    bad.write_text("ee:\n    type: se3\n    tracked: no_such_frame\n    weight: 1.0\n    kp: 1.0\n    mask: 111111\n")
    with pytest.raises(iw.IwbcError, match="no_such_frame"):
        iw.Problem("franka/pos_tracker.yaml", overrides={"tasks": str(bad)})
    bad.write_text("ee:\n    type: se3\n    tracked: panda_joint7\n    kp: 1.0\n    mask: 111111\n")
    with pytest.raises(iw.IwbcError, match="weight"):
        iw.Problem("franka/pos_tracker.yaml", overrides={"tasks": str(bad)})
    # humanoid controllers refuse a start posture whose CoM is not between the feet
    with pytest.raises(iw.IwbcError, match="CoM needs to be between the two feet"):
        iw.Problem("talos/talos_pos_tracker.yaml", overrides={"ref_config": "shifted"},
                   etc_path=_shifted_etc(tmp_path))


def _shifted_etc(tmp_path):
    import shutil
    dst = tmp_path / "etc"
    shutil.copytree(iw.ETC_PATH, dst)
    srdf = (dst / "talos" / "configurations.srdf").read_text()
    block = srdf[srdf.index("<group_state"):srdf.index("</group_state>") + len("</group_state>")]
    shifted = block.replace('name="inria_start"', 'name="shifted"').replace('name="torso_2_joint" value="0.0973496"', 'name="torso_2_joint" value="0.7"')
    (dst / "talos" / "configurations.srdf").write_text(srdf.replace("</robot>", shifted + "\n</robot>"))
    return str(dst)


def test_reference_yaml_dialect_parses(tmp_path):
    # the dialect of the reference's own files: comments, unquoted masks, multi-line flow lists, quoted strings
    text = '''CONTROLLER:
  name: pos-tracker
  solver: eiquadprog #qpmad
  base_path: "%s/franka"
  tasks: tasks.yaml
  urdf: franka.urdf
  configurations: configurations.srdf 
  ref_config: start
  floating_base: false
  floating_base_joint_name: ""
  dt: 0.001
  # a comment line
  closed_loop: false
  verbose: false
  mimic_dof_names : ["a",
                     "b"]
  check_model_collisions: false
''' % iw.ETC_PATH
    f = tmp_path / "c.yaml"
    f.write_text(text)
    P = iw.Problem(str(f))
    assert P.names("mimics") == ["a", "b"] and P.nv == 9


def test_set_task_weight(franka):
    P = iw.Problem("franka/pos_tracker.yaml")
    P.set_task_weight("posture", 0.25)
    assert P.struct().tasks[P.names("tasks").index("posture")].weight == 0.25
    with pytest.raises(iw.IwbcError):
        P.set_task_weight("nope", 1.0)


def test_no_gpu_fails_loudly(franka):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(iw.IwbcError, match="no CPU fallback"):
        iw.Batch(franka, 0, 16)
