"""URDF -> Mechanism with the reference's exact ordering rules (src/urdf/parse.jl:162-221).

Rules that decide the q/v/τ index order (get one wrong and every batched array is scrambled):
  * only DIRECT children <link>/<joint> of <robot> are read (parse.jl:184-185) -- transmissions
    contain nested <joint> tags that must be ignored;
  * the spanning tree is breadth-first from the unique root link, children in document order of the
    <joint> elements (parse.jl:187-206, graphs/spanning_tree.jl:45-83);
  * the root link is attached to the world by "<rootlink>_to_world", QuaternionFloating iff
    ``floating`` else Fixed (parse.jl:121-127);
  * fixed tree joints are then removed, the others keep their relative order (parse.jl:216-218);
  * <origin rpy> -> Rz(yaw) Ry(pitch) Rx(roll) (parse.jl:46-51); <inertia> is about the COM in the
    <inertial><origin> frame and is transformed into the link frame (parse.jl:104-112);
  * <dynamics damping>, <limit>, <mimic> do not enter the dynamics (parse.jl:74-95).
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from collections import deque
from typing import Dict, Optional

import numpy as np

from .joint_types import (Fixed, JointType, Planar, Prismatic, QuaternionFloating, Revolute)
from .mechanism import DEFAULT_GRAVITATIONAL_ACCELERATION, Joint, Mechanism, RigidBody
from .spatial import SpatialInertia, Transform3D, rot_rpy, rotation_between


def default_urdf_joint_types() -> Dict[str, type]:
    """parse.jl:6-15."""
    return {"revolute": Revolute, "continuous": Revolute, "prismatic": Prismatic,
            "floating": QuaternionFloating, "fixed": Fixed, "planar": Planar}


def _vec(e: Optional[ET.Element], name: str, default: str):
    s = default if e is None or e.get(name) is None else e.get(name)
    return np.array([float(x) for x in s.split()])


def _scalar(e: Optional[ET.Element], name: str, default: str = "0"):
    s = default if e is None or e.get(name) is None else e.get(name)
    return float(s)


def parse_pose(xml_pose: Optional[ET.Element]):
    """parse.jl:40-51."""
    if xml_pose is None:
        return np.eye(3), np.zeros(3)
    rpy = _vec(xml_pose, "rpy", "0 0 0")
    return rot_rpy(rpy[0], rpy[1], rpy[2]), _vec(xml_pose, "xyz", "0 0 0")


def _parse_joint_type(xml_joint: ET.Element, joint_types: Dict[str, type]) -> JointType:
    """parse.jl:53-72."""
    t = xml_joint.get("type")
    if t not in joint_types:
        raise ValueError(f"joint type {t} not recognized")
    cls = joint_types[t]
    if t in ("revolute", "continuous", "prismatic"):
        return cls(_vec(xml_joint.find("axis"), "xyz", "1 0 0"))
    if t in ("floating", "fixed"):
        return cls()
    if t == "planar":
        urdf_axis = _vec(xml_joint.find("axis"), "xyz", "1 0 0")
        R = rotation_between([0.0, 0.0, 1.0], urdf_axis)      # plane perpendicular to the URDF axis
        return cls(R @ np.array([1.0, 0, 0]), R @ np.array([0, 1.0, 0]))
    raise ValueError(f"joint type {t} not recognized")


def _parse_body(xml_link: ET.Element) -> RigidBody:
    """parse.jl:104-119: inertia about the COM in the <inertial><origin> frame, moved to the link frame;
    links without <inertial> get zero inertia."""
    xi = xml_link.find("inertial")
    if xi is None:
        inertia = SpatialInertia.zero()
    else:
        e = xi.find("inertia")
        ixx, ixy, ixz = _scalar(e, "ixx"), _scalar(e, "ixy"), _scalar(e, "ixz")
        iyy, iyz, izz = _scalar(e, "iyy"), _scalar(e, "iyz"), _scalar(e, "izz")
        moment = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        mass = _scalar(xi.find("mass"), "value", "0")
        rot, trans = parse_pose(xi.find("origin"))
        inertia = SpatialInertia(moment, np.zeros(3), mass).transform(Transform3D(rot, trans))
    return RigidBody(xml_link.get("name"), inertia)


def parse_urdf(filename: str, *, floating: bool = False, joint_types: Optional[Dict[str, type]] = None,
               root_joint_type: Optional[JointType] = None, remove_fixed_tree_joints: bool = True,
               gravity=DEFAULT_GRAVITATIONAL_ACCELERATION) -> Mechanism:
    """Mirror of ``parse_urdf(filename; floating, joint_types, root_joint_type, remove_fixed_tree_joints,
    gravity)`` (parse.jl:162-221).  ``scalar_type`` is not a parameter: the host model is always fp64 and
    the batch dtype is chosen per call (promotion rule mechanism_state.jl:179-182)."""
    jt = default_urdf_joint_types() if joint_types is None else joint_types
    if root_joint_type is None:
        root_joint_type = jt["floating" if floating else "fixed"]()
    if floating and not root_joint_type.isfloating:
        raise ValueError("Ambiguous input arguments: `floating` specified, but `root_joint_type` is not a "
                         "floating joint type.")                                    # parse.jl:177-179
    xroot = ET.parse(filename).getroot()
    if xroot.tag != "robot":
        raise ValueError("URDF root element must be <robot>")
    xml_links = xroot.findall("link")          # direct children only
    xml_joints = xroot.findall("joint")
    name_to_link = {l.get("name"): l for l in xml_links}

    out_edges = {n: [] for n in name_to_link}   # document order == add_edge! order
    has_parent = set()
    for xj in xml_joints:
        p = xj.find("parent").get("link")
        c = xj.find("child").get("link")
        out_edges[p].append(xj)
        has_parent.add(c)
    roots = [n for n in name_to_link if n not in has_parent]
    if len(roots) != 1:
        raise ValueError("Can only handle a single root")                          # parse.jl:204

    # breadth-first spanning tree, FIFO over edges (graphs/spanning_tree.jl:45-83 with next_edge = first)
    tree_edges = []
    visited = {roots[0]}
    frontier = deque(out_edges[roots[0]])
    while frontier:
        e = frontier.popleft()
        child = e.find("child").get("link")
        if child in visited:
            continue                             # a second path to an already-placed link (loop): skipped
        visited.add(child)
        tree_edges.append(e)
        frontier.extend(out_edges[child])
    if len(visited) != len(name_to_link):
        raise ValueError("Graph is not connected.")

    mech = Mechanism(RigidBody("world"), gravity=gravity)
    bodies = {}
    root_link = name_to_link[roots[0]]
    body = _parse_body(root_link)
    bodies[roots[0]] = body
    mech.attach(mech.root_body, body, Joint(f"{body.name}_to_world", root_joint_type))   # parse.jl:121-127
    for e in tree_edges:
        parent = bodies[e.find("parent").get("link")]
        child_name = e.find("child").get("link")
        joint = Joint(e.get("name"), _parse_joint_type(e, jt))
        rot, trans = parse_pose(e.find("origin"))
        body = _parse_body(name_to_link[child_name])
        bodies[child_name] = body
        mech.attach(parent, body, joint, joint_pose=Transform3D(rot, trans))             # parse.jl:129-140
    if remove_fixed_tree_joints:
        mech.remove_fixed_tree_joints()
    return mech
