#!/usr/bin/env python3
"""Derive the humanoid model tables from the reference MJCF and emit them as a C header + JSON.

Input : /root/reference/humanoid_symmetric_2.xml   (read-only reference file, never copied)
Output: imitation-learning-rl_b200/csrc/ilrl_model_data.h   (plain initialiser-list macros, no types)
        imitation-learning-rl_b200/data/model.json          (same numbers, for the Python host side)

Conventions restated from Bullet's MJCF importer (un-vendored third party, see DESIGN.md "Model"):
  * one link per <body>, inertial frame at the body-frame origin (no COM shift);
  * mass = 1000 kg/m^3 x geom volume (capsule = cylinder + sphere, sphere);
  * diagonal inertia = box inertia of the AABB of the body's geoms (btCompoundShape::calculateLocalInertia),
    from-to capsules being 2-sphere hulls (exact capsule AABB);
  * one hinge per <joint>, in document order (= DFS pre-order = pybullet link order = `ordered_joints`),
    several joints of a body compose in document order about their (shared) anchor;
  * <default> joint damping/armature/stiffness are not honoured; limits are (ranges are in degrees).
Reference call sites that fix the remaining numbers: humanoid.py:23 (power), :28-37 (motor order and gears).
"""
import json
import math
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF_XML = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/humanoid_symmetric_2.xml"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_H = os.path.join(ROOT, "imitation-learning-rl_b200", "csrc", "ilrl_model_data.h")
OUT_JSON = os.path.join(ROOT, "imitation-learning-rl_b200", "data", "model.json")

DENSITY = 1000.0

# humanoid.py:28-37 — motor (action slot) order and gears; power = 0.41 (humanoid.py:23)
MOTOR_NAMES = ["abdomen_z", "abdomen_y", "abdomen_x",
               "right_hip_x", "right_hip_z", "right_hip_y", "right_knee",
               "left_hip_x", "left_hip_z", "left_hip_y", "left_knee",
               "right_shoulder_x", "right_shoulder_y", "right_elbow",
               "left_shoulder_x", "left_shoulder_y", "left_elbow"]
MOTOR_POWER = [100, 100, 100, 100, 100, 300, 200, 100, 100, 300, 200, 75, 75, 75, 75, 75, 75]
POWER = 0.41

# low_level_env.py:86-137 — joint_map iteration order, CSV column names and the two weight tables
JOINT_MAP = [("right_knee", "rightKnee", 3, 1), ("right_hip_x", "rightHipX", 1, 1),
             ("right_hip_y", "rightHipY", 3, 1), ("right_hip_z", "rightHipZ", 1, 1),
             ("left_knee", "leftKnee", 3, 1), ("left_hip_x", "leftHipX", 1, 1),
             ("left_hip_y", "leftHipY", 3, 1), ("left_hip_z", "leftHipZ", 1, 1),
             ("right_shoulder_x", "rightShoulderX", 0.1, 0.1), ("right_shoulder_y", "rightShoulderY", 0.3, 0.1),
             ("right_elbow", "rightElbow", 0.3, 0.1), ("left_shoulder_x", "leftShoulderX", 0.1, 0.1),
             ("left_shoulder_y", "leftShoulderY", 0.3, 0.1), ("left_elbow", "leftElbow", 0.3, 0.1)]
CSV_COLS = ["rightHipX", "rightHipY", "rightHipZ", "rightKnee", "leftHipX", "leftHipY", "leftHipZ", "leftKnee",
            "rightShoulderX", "rightShoulderY", "rightElbow", "leftShoulderX", "leftShoulderY", "leftElbow"]
# contact-sphere priority: feet, shins, thighs, hands, forearms, rest (first MAX_CONTACTS active ones are kept)
CONTACT_PRIORITY = ["right_foot", "left_foot", "right_shin", "left_shin", "right_thigh", "left_thigh",
                    "right_hand", "left_hand", "right_lower_arm", "left_lower_arm", "pelvis", "lwaist",
                    "right_upper_arm", "left_upper_arm", "torso"]


def vec(s):
    return np.array([float(x) for x in s.split()], dtype=np.float64)


def main():
    root = ET.parse(REF_XML).getroot()
    assert root.find("compiler").get("angle") == "degree"
    bodies, joints, geoms = [], [], []

    def walk(el, parent):
        bi = len(bodies)
        q = vec(el.get("quat", "1 0 0 0"))
        q = q / np.linalg.norm(q)
        bodies.append(dict(name=el.get("name"), parent=parent, pos=vec(el.get("pos", "0 0 0")),
                           quat_xyzw=np.array([q[1], q[2], q[3], q[0]])))
        for ch in el:
            if ch.tag == "joint":
                ax = vec(ch.get("axis"))
                rng = vec(ch.get("range"))
                joints.append(dict(name=ch.get("name"), body=bi, anchor=vec(ch.get("pos", "0 0 0")),
                                   axis=ax / np.linalg.norm(ax), lo=math.radians(rng[0]), hi=math.radians(rng[1])))
            elif ch.tag == "geom":
                g = dict(name=ch.get("name"), body=bi, type=ch.get("type"), r=vec(ch.get("size"))[0])
                if ch.get("fromto") is not None:
                    ft = vec(ch.get("fromto"))
                    g["p0"], g["p1"] = ft[:3], ft[3:]
                else:
                    assert g["type"] == "sphere"
                    g["p0"] = g["p1"] = vec(ch.get("pos", "0 0 0"))
                geoms.append(g)
        for ch in el:
            if ch.tag == "body":
                walk(ch, bi)

    walk(root.find("worldbody").find("body"), -1)
    nb, nj = len(bodies), len(joints)
    assert nb == 15 and nj == 17

    # masses + AABB inertia per body
    for bi, b in enumerate(bodies):
        vol, lo, hi = 0.0, np.full(3, np.inf), np.full(3, -np.inf)
        for g in geoms:
            if g["body"] != bi:
                continue
            r = g["r"]
            h = np.linalg.norm(g["p1"] - g["p0"])
            vol += math.pi * r * r * h + 4.0 / 3.0 * math.pi * r ** 3
            for p in (g["p0"], g["p1"]):
                lo = np.minimum(lo, p - r)
                hi = np.maximum(hi, p + r)
        m = DENSITY * vol
        l = hi - lo
        b["mass"] = m
        b["inertia"] = m / 12.0 * np.array([l[1] ** 2 + l[2] ** 2, l[0] ** 2 + l[2] ** 2, l[0] ** 2 + l[1] ** 2])

    # joint tree: parent joint = last joint of the nearest ancestor body that has joints (or previous joint of the
    # same body); -1 = floating base (torso)
    last_joint_of_body = {}
    for ji, j in enumerate(joints):
        same = [k for k in range(ji) if joints[k]["body"] == j["body"]]
        if same:
            j["parent"] = same[-1]
        else:
            pb = bodies[j["body"]]["parent"]
            while pb >= 0 and pb not in last_joint_of_body:
                pb = bodies[pb]["parent"]
            j["parent"] = last_joint_of_body.get(pb, -1) if pb >= 0 else -1
        last_joint_of_body[j["body"]] = ji
    # the dynamic link a body is rigidly carried by = last joint of itself or of its nearest jointed ancestor
    for bi, b in enumerate(bodies):
        pb = bi
        while pb >= 0 and pb not in last_joint_of_body:
            pb = bodies[pb]["parent"]
        b["link"] = last_joint_of_body.get(pb, -1) if pb >= 0 else -1

    # composite rigid groups: a jointed (or root) body plus its joint-less descendants, expressed in the main frame
    comps = []
    for bi, b in enumerate(bodies):
        if bi != 0 and bi not in last_joint_of_body:
            continue
        members = [(bi, np.zeros(3))]
        stack = [(bi, np.zeros(3))]
        while stack:
            pi, off = stack.pop()
            for ci, c in enumerate(bodies):
                if c["parent"] == pi and ci not in last_joint_of_body:
                    assert np.allclose(c["quat_xyzw"], [0, 0, 0, 1])
                    members.append((ci, off + c["pos"]))
                    stack.append((ci, off + c["pos"]))
        m = sum(bodies[k]["mass"] for k, _ in members)
        com = sum(bodies[k]["mass"] * off for k, off in members) / m
        I = np.zeros((3, 3))
        for k, off in members:
            d = off - com
            I += np.diag(bodies[k]["inertia"]) + bodies[k]["mass"] * (d.dot(d) * np.eye(3) - np.outer(d, d))
        assert np.allclose(I, np.diag(np.diag(I)), atol=1e-12)
        comps.append(dict(body=bi, link=b["link"], mass=m, com=com, inertia=np.diag(I),
                          members=[k for k, _ in members]))
    assert len(comps) == 11

    # contact spheres (capsule end points / spheres), priority ordered
    spheres = []
    for bname in CONTACT_PRIORITY:
        bi = [k for k, b in enumerate(bodies) if b["name"] == bname][0]
        for g in geoms:
            if g["body"] != bi:
                continue
            pts = [g["p0"]] if np.allclose(g["p0"], g["p1"]) else [g["p0"], g["p1"]]
            # lower end of limb capsules first (more likely to touch)
            pts = sorted(pts, key=lambda p: p[2])
            for p in pts:
                spheres.append(dict(body=bi, link=bodies[bi]["link"], c=p, r=g["r"]))
    assert len(spheres) == 29

    # self-collision pairs (Bullet: URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS): every pair of geoms whose bodies
    # are not ancestor-related, in geom order; each geom by the contact-sphere indices of its axis end points (p0, p1)
    def is_anc(a, b):
        while b >= 0:
            if b == a:
                return True
            b = bodies[b]["parent"]
        return False

    def sph_of(g, p):
        k = [i for i, sp in enumerate(spheres) if sp["body"] == g["body"] and np.allclose(sp["c"], p) and sp["r"] == g["r"]]
        assert len(k) == 1, (g["name"], k)
        return k[0]

    self_pairs, self_reach = [], []
    for ia, ga in enumerate(geoms):
        for gb in geoms[ia + 1:]:
            if is_anc(ga["body"], gb["body"]) or is_anc(gb["body"], ga["body"]):
                continue
            self_pairs.append((sph_of(ga, ga["p0"]), sph_of(ga, ga["p1"]), sph_of(gb, gb["p0"]), sph_of(gb, gb["p1"])))
            # no contact is possible while the axis midpoints are farther apart than this (half lengths + radii + the
            # contact breaking distance 0.02, 1 mm of slack): the broad phase of the pair tests
            self_reach.append(0.5 * np.linalg.norm(ga["p1"] - ga["p0"]) + 0.5 * np.linalg.norm(gb["p1"] - gb["p0"]) +
                              ga["r"] + gb["r"] + 0.02 + 0.001)
    assert len(self_pairs) == 66

    jidx = {j["name"]: k for k, j in enumerate(joints)}
    motor_joint = [jidx[n] for n in MOTOR_NAMES]
    gear_motor = [p * POWER for p in MOTOR_POWER]
    map_joint = [jidx[a] for a, _, _, _ in JOINT_MAP]
    map_col = [CSV_COLS.index(c) for _, c, _, _ in JOINT_MAP]
    map_w = [w for _, _, w, _ in JOINT_MAP]
    map_wv = [w for _, _, _, w in JOINT_MAP]

    total_mass = sum(b["mass"] for b in bodies)

    def arr(x):
        return np.asarray(x).reshape(-1).tolist()

    model = dict(
        nb=nb, nj=nj, nc=len(comps), ns=len(spheres), total_mass=total_mass,
        body_name=[b["name"] for b in bodies], body_parent=[b["parent"] for b in bodies],
        body_link=[b["link"] for b in bodies],
        body_pos=arr([b["pos"] for b in bodies]), body_quat=arr([b["quat_xyzw"] for b in bodies]),
        body_mass=[b["mass"] for b in bodies], body_inertia=arr([b["inertia"] for b in bodies]),
        joint_name=[j["name"] for j in joints], joint_body=[j["body"] for j in joints],
        joint_parent=[j["parent"] for j in joints], joint_anchor=arr([j["anchor"] for j in joints]),
        joint_axis=arr([j["axis"] for j in joints]), joint_lo=[j["lo"] for j in joints],
        joint_hi=[j["hi"] for j in joints],
        comp_body=[c["body"] for c in comps], comp_link=[c["link"] for c in comps],
        comp_mass=[c["mass"] for c in comps], comp_com=arr([c["com"] for c in comps]),
        comp_inertia=arr([c["inertia"] for c in comps]),
        sphere_body=[s["body"] for s in spheres], sphere_link=[s["link"] for s in spheres],
        sphere_c=arr([s["c"] for s in spheres]), sphere_r=[s["r"] for s in spheres],
        # every collision geom as a capsule (a sphere is a zero-length one), body frame: self-collision pairs
        ng=len(geoms), geom_name=[g["name"] for g in geoms], geom_body=[g["body"] for g in geoms],
        geom_p0=arr([g["p0"] for g in geoms]), geom_p1=arr([g["p1"] for g in geoms]), geom_r=[g["r"] for g in geoms],
        nself=len(self_pairs), self_a0=[p[0] for p in self_pairs], self_a1=[p[1] for p in self_pairs],
        self_b0=[p[2] for p in self_pairs], self_b1=[p[3] for p in self_pairs], self_reach=self_reach,
        motor_names=MOTOR_NAMES, motor_joint=motor_joint, motor_gear=gear_motor,
        map_joint=map_joint, map_col=map_col, map_w=map_w, map_wv=map_wv, csv_cols=CSV_COLS,
    )
    os.makedirs(os.path.dirname(OUT_JSON), exist_ok=True)
    with open(OUT_JSON, "w") as f:
        json.dump(model, f, indent=1)

    def cl(x, fmt="%.17g"):
        return "{" + ", ".join(fmt % v if not isinstance(v, int) else "%d" % v for v in x) + "}"

    L = ["/* GENERATED by tools/gen_model.py from the reference MJCF (humanoid_symmetric_2.xml) - do not edit.",
         " * Plain initialiser lists: every consumer declares its own typed array (double in oracle/, float in csrc/).",
         " * Index spaces: body 0..14 (document order), joint 0..16 (= pybullet `ordered_joints` order),",
         " * comp 0..10 (rigid groups: a jointed body + its joint-less children), sphere 0..28 (ground-contact",
         " * candidates, priority ordered), motor 0..16 (action slots, humanoid.py:28-37), map 0..13 (joint_map order,",
         " * low_level_env.py:86-101). link == joint index that carries the body, -1 = floating base (torso). */",
         "#ifndef ILRL_MODEL_DATA_H", "#define ILRL_MODEL_DATA_H",
         "#define ILRL_NB 15", "#define ILRL_NJ 17", "#define ILRL_NC 11", "#define ILRL_NS 29", "#define ILRL_NMAP 14",
         "#define ILRL_NG %d" % len(geoms), "#define ILRL_NSELF %d" % len(self_pairs), "#define ILRL_TOTAL_MASS %.17g" % total_mass]
    for key in ["body_parent", "body_link", "body_pos", "body_quat", "body_mass", "body_inertia", "joint_body",
                "joint_parent", "joint_anchor", "joint_axis", "joint_lo", "joint_hi", "comp_body", "comp_link",
                "comp_mass", "comp_com", "comp_inertia", "sphere_body", "sphere_link", "sphere_c", "sphere_r",
                "geom_body", "geom_p0", "geom_p1", "geom_r", "self_a0", "self_a1", "self_b0", "self_b1", "self_reach", "motor_joint", "motor_gear", "map_joint", "map_col", "map_w", "map_wv"]:
        L.append("#define ILRL_%s %s" % (key.upper(), cl(model[key])))
    L.append("#endif")
    os.makedirs(os.path.dirname(OUT_H), exist_ok=True)
    with open(OUT_H, "w") as f:
        f.write("\n".join(L) + "\n")
    print("bodies:", [(b["name"], round(b["mass"], 4)) for b in bodies])
    print("total mass %.4f" % total_mass)
    print("joint parents:", model["joint_parent"])
    print("body link:", model["body_link"])
    print("comps:", [(bodies[c["body"]]["name"], round(c["mass"], 4), c["com"].round(4).tolist(),
                      c["inertia"].round(5).tolist()) for c in comps])


if __name__ == "__main__":
    main()
