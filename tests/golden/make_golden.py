"""Generates the golden fixtures under tests/golden/ from the reference tree (run in the build
container where /root/reference is mounted; the GPU box only sees the committed .npz/.json files).

For every deck that pins the hot path with a CLOAD-only load case:
  * the mesh as parsed by frontistr_b200.mshio.read_msh (arrays only, no reference source text)
  * the !BOUNDARY / !CLOAD / !SOLVER cards
  * the nodal displacement max/min lines of '##### Global Summary' from <deck>_correct.log
    (the quantities the reference's own harness compares, examples/test_FrontISTR.rb:148-237)
Also stores the 2-rank halo-table fixture examples/static/refine/cube/P02/sample_P02.{0,1}.
"""
import json
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from frontistr_b200 import mshio  # noqa: E402

REF = os.environ.get("FISTR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

DECKS = [
    ("A361", "examples/static/exA/A361.msh", "examples/static/exA/A300.cnt", "examples/static/exA/A361_correct.log"),
    ("A342", "examples/static/exA/A342.msh", "examples/static/exA/A300.cnt", "examples/static/exA/A342_correct.log"),
    ("G361", "examples/static/exG/A361.msh", "examples/static/exG/G300.cnt", "examples/static/exG/A361_correct.log"),
    ("G342", "examples/static/exG/A342.msh", "examples/static/exG/G300.cnt", "examples/static/exG/A342_correct.log"),
]


# geometrically nonlinear decks (examples/static/exI, `!STATIC, TYPE=NLGEOM`, `!STEP, SUBSTEPS=10`): same meshes
# as exA, so only the cards and the golden summary are stored; "mesh" names the fixture to load
NL_DECKS = [
    ("I361", "A361", "examples/static/exI/I300.cnt", "examples/static/exI/A361_correct.log"),
    ("I342", "A342", "examples/static/exI/I300.cnt", "examples/static/exI/A342_correct.log"),
]


# volume-load decks (examples/static/exC BZ, exD GRAV, exE CENT): !DLOAD on EGRP=ALL; density from the mesh's ITEM=2
DLOAD_DECKS = [
    ("C361", "examples/static/exC/C361.msh", "examples/static/exC/C300.cnt", "examples/static/exC/C361_correct.log"),
    ("C342", "examples/static/exC/C342.msh", "examples/static/exC/C300.cnt", "examples/static/exC/C342_correct.log"),
    ("D361", "examples/static/exD/D361.msh", "examples/static/exD/D300.cnt", "examples/static/exD/D361_correct.log"),
    ("D342", "examples/static/exD/D342.msh", "examples/static/exD/D300.cnt", "examples/static/exD/D342_correct.log"),
    ("E361", "examples/static/exE/E361.msh", "examples/static/exE/E300.cnt", "examples/static/exE/E361_correct.log"),
    ("E342", "examples/static/exE/E342.msh", "examples/static/exE/E300.cnt", "examples/static/exE/E342_correct.log"),
]
LTYPE = {"BX": 1, "BY": 2, "BZ": 3, "GRAV": 4, "CENT": 5}
# face-pressure decks (examples/static/exB): !DLOAD <EGRP>, P<face>, value
PRESSURE_DECKS = [
    ("B361", "examples/static/exB/B361.msh", "examples/static/exB/B361.cnt", "examples/static/exB/B361_correct.log"),
    ("B342", "examples/static/exB/B342.msh", "examples/static/exB/B342.cnt", "examples/static/exB/B342_correct.log"),
]


def element_ids(mtxt):
    """Element ids in file order (first integer of every data line of the !ELEMENT blocks)."""
    ids, inside = [], False
    for line in mtxt.splitlines():
        t = line.strip()
        if t.startswith("!"):
            inside = t.upper().startswith("!ELEMENT")
            continue
        if inside and t and not t.startswith("#"):
            ids.append(int(t.replace(",", " ").split()[0]))
    return ids


def egroup(mtxt, name):
    m = re.search(r"!EGROUP\s*,\s*EGRP\s*=\s*%s\s*(,\s*GENERATE)?\s*\n((?:\s*[-0-9][^\n]*\n)+)" % re.escape(name), mtxt)
    out = []
    for line in m.group(2).strip().splitlines():
        v = [int(x) for x in line.replace(",", " ").split()]
        out += list(range(v[0], v[1] + 1, v[2])) if m.group(1) else v
    return out


def parse_log(path, which=1):
    """Displacement max/min of the `which`-th '##### Global Summary' block (1-based; NL logs have one per sub-step)."""
    out = {}
    with open(path) as f:
        txt = f.read()
    blk = txt.split("##### Global Summary")[which]
    for line in blk.splitlines():
        m = re.match(r"\s*//(U[123])\s+(\S+)\s+(\S+)", line)
        if m:
            out[m.group(1)] = [float(m.group(2)), float(m.group(3))]
        if line.strip().startswith("##### @Element"):
            break
    return out


def parse_log_stress(path, which=1):
    """The strain / stress / Mises max/min lines of the `which`-th '##### Global Summary' block: nodal values
    (fstrSOLID%STRAIN/STRESS/MISES) and, after '##### @Element', elemental ones (ESTRAIN/ESTRESS/EMISES) — the
    output of fstr_UpdateNewton + fstr_NodalStress3D, compared by the reference's harness like U1-U3."""
    with open(path) as f:
        txt = f.read()
    blk = txt.split("##### Global Summary")[which]
    out = {"node": {}, "elem": {}}
    where = "node"
    for line in blk.splitlines()[1:]:
        if line.strip().startswith("##### @Element"):
            where = "elem"
            continue
        m = re.match(r"\s*//([ES]\d\d|SMS)\s+(\S+)\s+(\S+)\s*$", line)
        if m:
            out[where][m.group(1)] = [float(m.group(2)), float(m.group(3))]
        elif line.strip().startswith("####") or line.strip().startswith("#####"):
            break
    return out


def main():
    meta = {}
    for name, msh, cnt, log in DECKS:
        m = mshio.read_msh(os.path.join(REF, msh))
        c = mshio.read_cnt(os.path.join(REF, cnt))
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            node=m.node, elem_type_index=m.elem_type_index, elem_type_item=m.elem_type_item,
                            elem_node_index=m.elem_node_index, elem_node_item=m.elem_node_item,
                            global_node_ID=m.global_node_ID,
                            **{"ngrp_" + k: v for k, v in m.node_group.items()})
        meta[name] = dict(source=dict(msh=msh, cnt=cnt, log=log), E=m.sections[0]["E"], nu=m.sections[0]["nu"],
                          boundary=c["boundary"], cload=c["cload"], solver=[c["solver"][0], c["solver"][1]],
                          golden=parse_log(os.path.join(REF, log)), tolerance_abs=1.0e-4)
    for name, msh, cnt, log in DLOAD_DECKS:
        m = mshio.read_msh(os.path.join(REF, msh))
        c = mshio.read_cnt(os.path.join(REF, cnt))
        with open(os.path.join(REF, msh)) as f:
            mtxt = f.read()
        with open(os.path.join(REF, cnt)) as f:
            ctxt = f.read()
        rho = float(re.search(r"!ITEM\s*=\s*2[^\n]*\n\s*([^\s,]+)", mtxt).group(1))
        dl = re.search(r"!DLOAD[^\n]*\n\s*([^\n]+)", ctxt).group(1).replace(",", " ").split()
        assert dl[0] == "ALL" and dl[1] in LTYPE
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            node=m.node, elem_type_index=m.elem_type_index, elem_type_item=m.elem_type_item,
                            elem_node_index=m.elem_node_index, elem_node_item=m.elem_node_item,
                            global_node_ID=m.global_node_ID, **{"ngrp_" + k: v for k, v in m.node_group.items()})
        meta[name] = dict(source=dict(msh=msh, cnt=cnt, log=log), E=m.sections[0]["E"], nu=m.sections[0]["nu"], rho=rho,
                          boundary=c["boundary"], cload=[], dload=dict(ltype=LTYPE[dl[1]], params=[float(x) for x in dl[2:]]),
                          solver=[c["solver"][0], c["solver"][1]], golden=parse_log(os.path.join(REF, log)),
                          tolerance_abs=1.0e-4)
    for name, msh, cnt, log in PRESSURE_DECKS:
        m = mshio.read_msh(os.path.join(REF, msh))
        c = mshio.read_cnt(os.path.join(REF, cnt))
        with open(os.path.join(REF, msh)) as f:
            mtxt = f.read()
        with open(os.path.join(REF, cnt)) as f:
            ctxt = f.read()
        dl = re.search(r"!DLOAD[^\n]*\n\s*([^\n]+)", ctxt).group(1).replace(",", " ").split()
        ids = element_ids(mtxt)
        assert len(ids) == m.n_elem and re.fullmatch(r"P[1-6]", dl[1])
        pos = {eid: k for k, eid in enumerate(ids)}
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            node=m.node, elem_type_index=m.elem_type_index, elem_type_item=m.elem_type_item,
                            elem_node_index=m.elem_node_index, elem_node_item=m.elem_node_item,
                            global_node_ID=m.global_node_ID, **{"ngrp_" + k: v for k, v in m.node_group.items()})
        meta[name] = dict(source=dict(msh=msh, cnt=cnt, log=log), E=m.sections[0]["E"], nu=m.sections[0]["nu"],
                          boundary=c["boundary"], cload=[],
                          pressure=dict(face=int(dl[1][1]), value=float(dl[2]), elems=[pos[e] for e in egroup(mtxt, dl[0])]),
                          solver=[c["solver"][0], c["solver"][1]], golden=parse_log(os.path.join(REF, log)),
                          tolerance_abs=1.0e-4)
    # thermal-stress decks (examples/static/exF): !TEMPERATURE ALL, 120 ; !REFTEMP 20 ; initial temperature 20 ; ITEM=3 = alpha
    for name in ("F361", "F342"):
        msh = "examples/static/exF/%s.msh" % name
        cnt, log = "examples/static/exF/F300.cnt", "examples/static/exF/%s_correct.log" % name
        m = mshio.read_msh(os.path.join(REF, msh))
        c = mshio.read_cnt(os.path.join(REF, cnt))
        with open(os.path.join(REF, msh)) as f:
            mtxt = f.read()
        with open(os.path.join(REF, cnt)) as f:
            ctxt = f.read()
        alp = float(re.search(r"!ITEM\s*=\s*3[^\n]*\n\s*([^\s,]+)", mtxt).group(1))
        t0 = float(re.search(r"!INITIAL CONDITION\s*,\s*TYPE\s*=\s*TEMPERATURE\s*\n\s*ALL\s*,\s*([^\s,]+)", mtxt).group(1))
        tt = float(re.search(r"!TEMPERATURE[^\n]*\n\s*ALL\s*,\s*([^\s,]+)", ctxt).group(1))
        ref = float(re.search(r"!REFTEMP[^\n]*\n\s*([^\s,]+)", ctxt).group(1))
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            node=m.node, elem_type_index=m.elem_type_index, elem_type_item=m.elem_type_item,
                            elem_node_index=m.elem_node_index, elem_node_item=m.elem_node_item,
                            global_node_ID=m.global_node_ID, **{"ngrp_" + k: v for k, v in m.node_group.items()})
        meta[name] = dict(source=dict(msh=msh, cnt=cnt, log=log), E=m.sections[0]["E"], nu=m.sections[0]["nu"],
                          boundary=c["boundary"], cload=[], thermal=dict(alpha=alp, temperature=tt, initial=t0, reftemp=ref),
                          solver=[c["solver"][0], c["solver"][1]], golden=parse_log(os.path.join(REF, log)),
                          tolerance_abs=1.0e-4)
    for name, mesh_name, cnt, log in NL_DECKS:
        with open(os.path.join(REF, cnt)) as f:
            txt = f.read()
        assert re.search(r"!STATIC\s*,\s*TYPE\s*=\s*NLGEOM", txt)
        c = mshio.read_cnt(os.path.join(REF, cnt))
        E, nu = [float(x) for x in re.search(r"!ELASTIC\s*\n\s*([^\n]+)", txt).group(1).replace(",", " ").split()[:2]]
        meta[name] = dict(source=dict(mesh=mesh_name, cnt=cnt, log=log), mesh=mesh_name, E=E, nu=nu,
                          boundary=c["boundary"], cload=c["cload"], solver=[c["solver"][0], c["solver"][1]],
                          nlgeom=dict(substeps=int(re.search(r"SUBSTEPS\s*=\s*(\d+)", txt).group(1)),
                                      converg=1.0e-3, max_iter=50),   # defaults of lib/m_step.f90:69,77
                          golden_steps=[parse_log(os.path.join(REF, log), k) for k in range(1, 11)],
                          tolerance_abs=1.0e-4)
    # spring_boundary/SB1: NLSTATIC, 2 C3D8 (B-bar), !SPRING, CG + SSOR with iterPREmax=2, 5 sub-steps, golden log per step
    m = mshio.read_msh(os.path.join(REF, "examples/static/spring_boundary/SB.msh"))
    c = mshio.read_cnt(os.path.join(REF, "examples/static/spring_boundary/SB1.cnt"))
    np.savez_compressed(os.path.join(OUT, "SB.npz"),
                        node=m.node, elem_type_index=m.elem_type_index, elem_type_item=m.elem_type_item,
                        elem_node_index=m.elem_node_index, elem_node_item=m.elem_node_item,
                        global_node_ID=m.global_node_ID, **{"ngrp_" + k: v for k, v in m.node_group.items()})
    sb_log = os.path.join(REF, "examples/static/spring_boundary/SB_correct.log")
    meta["SB1"] = dict(source=dict(msh="examples/static/spring_boundary/SB.msh", cnt="examples/static/spring_boundary/SB1.cnt",
                                   log="examples/static/spring_boundary/SB_correct.log"),
                       mesh="SB", E=100.0, nu=0.3, boundary=c["boundary"], cload=c["cload"], spring=[["SET-2", 2, 1.0]],
                       solver=[c["solver"][0], c["solver"][1]], nlgeom=dict(substeps=5, converg=1.0e-6, max_iter=15),
                       golden_steps=[parse_log(sb_log, k) for k in range(1, 6)], golden={}, tolerance_abs=1.0e-4)
    # tutorial/03_hyperelastic_cylinder (BASELINE.json configs[4]): mesh + cards only — the reference ships no golden
    # log for it, so this deck is a run-only fixture (Newton convergence, prescribed displacement, symmetry planes)
    m = mshio.read_msh(os.path.join(REF, "tutorial/03_hyperelastic_cylinder/cylinder.msh"))
    c = mshio.read_cnt(os.path.join(REF, "tutorial/03_hyperelastic_cylinder/cylinder.cnt"))
    np.savez_compressed(os.path.join(OUT, "cylinder.npz"),
                        node=m.node, elem_type_index=m.elem_type_index, elem_type_item=m.elem_type_item,
                        elem_node_index=m.elem_node_index, elem_node_item=m.elem_node_item,
                        global_node_ID=m.global_node_ID, **{"ngrp_" + k: v for k, v in m.node_group.items()})
    meta["cylinder"] = dict(source=dict(msh="tutorial/03_hyperelastic_cylinder/cylinder.msh",
                                        cnt="tutorial/03_hyperelastic_cylinder/cylinder.cnt", log=None),
                            E=1.0, nu=0.3, mooney_rivlin=[0.1486, 0.4849, 0.0789], boundary=c["boundary"], cload=[],
                            solver=[c["solver"][0], c["solver"][1]], nlgeom=dict(substeps=5, converg=1.0e-5, max_iter=50),
                            golden={}, tolerance_abs=None)
    # 2-rank halo fixture
    for r in (0, 1):
        d = mshio.read_dist_mesh(os.path.join(REF, "examples/static/refine/cube/P02/sample_P02.%d" % r))
        np.savez_compressed(os.path.join(OUT, "sample_P02_%d.npz" % r),
                            n_node=d.n_node, nn_internal=d.nn_internal, node=d.node, node_ID=d.node_ID,
                            global_node_ID=d.global_node_ID, elem_type_index=d.elem_type_index,
                            elem_type_item=d.elem_type_item, elem_node_index=d.elem_node_index,
                            elem_node_item=d.elem_node_item, neighbor_pe=d.neighbor_pe,
                            import_index=d.import_index, import_item=d.import_item,
                            export_index=d.export_index, export_item=d.export_item, my_rank=d.my_rank, PETOT=d.PETOT)
    # strain / stress / Mises lines of the linear decks (exA-exG)
    for name, d in meta.items():
        log = d["source"].get("log")
        if log and "nlgeom" not in d:
            d["golden_stress"] = parse_log_stress(os.path.join(REF, log))
            assert len(d["golden_stress"]["node"]) == 13 and len(d["golden_stress"]["elem"]) == 13, name
    # ... and of every sub-step of the nonlinear decks (Green-Lagrange strain / 2nd Piola-Kirchhoff stress of the
    # total-Lagrangian update, through the same fstr_NodalStress3D)
    for name in ("I361", "I342", "SB1"):
        d = meta[name]
        d["golden_stress_steps"] = [parse_log_stress(os.path.join(REF, d["source"]["log"]), k)
                                    for k in range(1, d["nlgeom"]["substeps"] + 1)]
        assert all(len(x["node"]) == 13 and len(x["elem"]) == 13 for x in d["golden_stress_steps"]), name
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
