"""Boundary hygiene (CPU, this container only): the C++ host mirror (cddp-cpp_b200/host/include/cddp_b200/cddp.hpp) is
diffed against the reference's own public headers, so that the shim of INTEGRATION.md keeps compiling against
astomodynamics/cddp-cpp:
  * every option member the mirror declares exists in the reference's struct of the same name with the same default
    (include/cddp-cpp/cddp_core/options.hpp, boxqp.hpp),
  * every public method of the mirrored classes that the solver plugin calls exists in the reference class with the same
    parameter list (cddp_core.hpp, dynamical_system.hpp, objective.hpp, constraint.hpp),
  * the mirror's additions to the reference surface are exactly the documented ones.
Skipped where /root/reference is absent (the GPU box)."""
import os
import re

import pytest

REF = "/root/reference/include/cddp-cpp/cddp_core"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIRROR = os.path.join(ROOT, "cddp-cpp_b200", "host", "include", "cddp_b200", "cddp.hpp")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def strip_comments(txt):
    txt = re.sub(r"/\*[\s\S]*?\*/", "", txt)
    return re.sub(r"//.*", "", txt)


def body_of(txt, kind, name):
    """text between the braces of `struct|class name ... { ... };` (first definition)"""
    m = re.search(rf"\b{kind}\s+{name}\b[^;{{]*{{", txt)
    assert m, f"{kind} {name} not found"
    depth, i = 1, m.end()
    while depth:
        depth += {"{": 1, "}": -1}.get(txt[i], 0)
        i += 1
    return txt[m.end(): i - 1]


def members_with_defaults(body):
    """{member: normalised default} for `type name = value;` declarations at depth 0 of a struct body"""
    out, depth, stmt = {}, 0, ""
    for ch in body:
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
        if depth == 0:
            stmt += ch
            if ch == ";":
                m = re.match(r"\s*([\w:<>\s]+?)\s+(\w+)\s*=\s*([^;]+);", stmt.replace("\n", " "))
                if m:
                    out[m.group(2)] = re.sub(r"\s+", "", m.group(3))
                stmt = ""
    return out


def num(v):
    v = v.replace("BarrierStrategy::", "")
    try:
        return float(v)
    except ValueError:
        return v.strip('"')


OPTION_STRUCTS = [("LineSearchOptions", "options.hpp"), ("RegularizationOptions", "options.hpp"), ("BoxQPOptions", "boxqp.hpp"),
                  ("SolverSpecificFilterOptions", "options.hpp"), ("SolverSpecificBarrierOptions", "options.hpp"),
                  ("IPDDPAlgorithmOptions", "options.hpp"), ("CDDPOptions", "options.hpp")]


@pytest.mark.parametrize("name,header", OPTION_STRUCTS)
def test_option_members_and_defaults_match_the_reference(name, header):
    ref = strip_comments(open(os.path.join(REF, header)).read())
    mir = strip_comments(open(MIRROR).read())
    rm, mm = members_with_defaults(body_of(ref, "struct", name)), members_with_defaults(body_of(mir, "struct", name))
    assert mm, name
    for member, default in mm.items():
        if name == "IPDDPAlgorithmOptions" and member in ("barrier",):
            continue
        assert member in rm, f"{name}::{member} is not a member of the reference's {name}"
        assert num(default) == num(rm[member]), f"{name}::{member}: mirror default {default}, reference {rm[member]}"


def public_methods(body):
    """{name: [normalised parameter lists]} of the function declarations in the public sections of a class body"""
    out, public = {}, False
    flat, depth, cur = [], 0, ""
    for ch in body:  # drop inline function bodies
        if ch == "{":
            depth += 1
            if depth == 1:
                cur += ";"
            continue
        if ch == "}":
            depth -= 1
            continue
        if depth == 0:
            cur += ch
    for stmt in re.split(r";", cur):
        s = " ".join(stmt.split())
        for label in re.findall(r"\b(public|private|protected)\s*:", s):
            public = label == "public"
        s = re.sub(r"\b(public|private|protected)\s*:", "", s).strip()
        if not public or "(" not in s:
            continue
        m = re.match(r"(?:template\s*<[^>]*>\s*)?(?:virtual\s+|static\s+|explicit\s+|inline\s+)*[\w:<>,&\*\s]*?\b(~?\w+)\s*\(", s)
        if not m or m.group(1) in ("if", "for", "while", "return"):
            continue
        depth, j = 1, m.end()  # the parameter list ends at the parenthesis that matches the first one
        while depth and j < len(s):
            depth += {"(": 1, ")": -1}.get(s[j], 0)
            j += 1
        params = []
        for prm in split_params(s[m.end(): j - 1]):
            prm = re.sub(r"=.*", "", prm).strip()
            m2 = re.match(r"(.*[\w>&\*])\s*\b(\w+)$", prm)  # drop the parameter name (if what is left is still a type)
            if m2 and re.search(r"[\w>]", m2.group(1)) and m2.group(1).strip() not in ("const", "unsigned"):
                prm = m2.group(1)
            params.append(re.sub(r"\s+", "", prm.replace("Eigen::", "").replace("std::", "")))
        out.setdefault(m.group(1), []).append(tuple(p for p in params if p))
    return out


def split_params(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        depth += {"<": 1, "(": 1, ">": -1, ")": -1}.get(ch, 0)
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


# class -> (reference header, methods of the mirror that are NOT in the reference: the documented additions)
CLASSES = {
    "ISolverAlgorithm": ("cddp_core.hpp", set()),
    "DynamicalSystem": ("dynamical_system.hpp", {"getDeviceModel"}),
    "CDDP": ("cddp_core.hpp", set()),
}


@pytest.mark.parametrize("cls", sorted(CLASSES))
def test_public_methods_exist_in_the_reference_with_the_same_parameters(cls):
    header, additions = CLASSES[cls]
    ref = public_methods(body_of(strip_comments(open(os.path.join(REF, header)).read()), "class", cls))
    mir = public_methods(body_of(strip_comments(open(MIRROR).read()), "class", cls))
    assert len(mir) >= 3, (cls, mir)
    extra = {m for m in mir if m not in ref and not m.startswith("~")}
    assert extra == additions, f"{cls}: mirror methods absent from the reference: {sorted(extra - additions)}; stale additions: {sorted(additions - extra)}"
    for name, overloads in mir.items():
        if name in additions or name.startswith("~"):
            continue
        for params in overloads:
            assert any(len(params) == len(r) and all(a == b for a, b in zip(params, r)) for r in ref[name]), \
                f"{cls}::{name}{params} has no overload with these parameter types in the reference: {ref[name]}"


def test_solution_struct_fields_match_the_reference():
    ref = strip_comments(open(os.path.join(REF, "cddp_core.hpp")).read())
    mir = strip_comments(open(MIRROR).read())

    def fields(body):
        return set(re.findall(r"\b(\w+)\s*(?:=\s*[^;{]+|\{[^}]*\})?\s*;", re.sub(r"\b(struct|class)\s+\w+\s*{[\s\S]*?};", "", body)))

    rf, mf = fields(body_of(ref, "struct", "CDDPSolution")), fields(body_of(mir, "struct", "CDDPSolution"))
    must = {"solver_name", "status_message", "iterations_completed", "solve_time_ms", "final_objective", "final_step_length",
            "final_regularization", "time_points", "state_trajectory", "control_trajectory", "feedback_gains"}
    assert must <= rf, must - rf
    assert must <= mf, must - mf
