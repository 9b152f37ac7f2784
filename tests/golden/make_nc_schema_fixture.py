"""Extracts the definitions of the reference's output database from its Fortran and stores them as
tests/golden/nc_schema_reference.json: every nf90_def_grp / nf90_def_dim / nf90_def_var of
nc_define_outputfile (SOLVER/nc_routines.F90:829-1492), the snapshot variable lists per dump type
(:943-1045) and the names of the global attributes (SOLVER/parameters.F90:1480-1552).

    python tests/golden/make_nc_schema_fixture.py      (needs /root/reference; the fixture is committed)
"""
import json
import os
import re

REF = "/root/reference/SOLVER"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(os.path.join(REF, "nc_routines.F90")).read()
    body = src[src.index("subroutine nc_define_outputfile"):src.index("end subroutine nc_define_outputfile")]
    body = re.sub(r"&\s*\n\s*&?", " ", body)
    grp_of = {"ncid_out": "", "ncid_recout": "Seismograms", "ncid_snapout": "Snapshots", "ncid_surfout": "Surface",
              "ncid_meshout": "Mesh"}
    groups, dims, dimid, variables = [], {}, {}, []
    calls = []
    for m in re.finditer(r"nf90_def_(grp|dim|var)\s*\(", body):
        depth, k = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(body[k], 0)
            k += 1
        calls.append((m.group(1), body[m.end():k - 1]))
    for kind, args in calls:
        args = re.sub(r"\s+", " ", args)
        parts, depth, cur = [], 0, ""
        for ch in args:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += ch
        parts.append(cur.strip())
        kw, pos = {}, []
        for a in parts:
            mm = re.match(r"^(\w+)\s*=\s*(.*)$", a)
            if mm and not a.startswith(("'", '"')):
                kw[mm.group(1)] = mm.group(2).strip()
            else:
                pos.append(a)
        if kind == "grp":
            groups.append(pos[1].strip("\"'"))
        elif kind == "dim":
            ncid = (kw.get("ncid") or pos[0]).strip()
            name = (kw.get("name") or pos[1]).strip().strip("\"'")
            length = (kw.get("len") or pos[2]).strip()
            did = (kw.get("dimid") or pos[3]).strip()
            dims[name] = {"group": grp_of[ncid], "len": length}
            dimid[did] = name
        else:
            ncid = (kw.get("ncid") or pos[0]).strip()
            name = (kw.get("name") or pos[1]).strip().strip("\"'")
            xtype = (kw.get("xtype") or pos[2]).strip()
            d = (kw.get("dimids") or pos[3]).strip().strip("[]")
            variables.append({"group": grp_of[ncid], "name": name, "xtype": xtype,
                              "dimids_fortran": [x.strip() for x in d.split(",")]})
    for v in variables:
        v["dims_fortran"] = [dimid[x] for x in v["dimids_fortran"]]
        del v["dimids_fortran"]
    # snapshot variable lists
    lists = {}
    for m in re.finditer(r"case \('(\w+)'\)(.*?)(?=case \(|end select)", body, re.S):
        names = re.findall(r"nc_varnamelist = \[(.*?)\]", m.group(2), re.S)
        if names:
            lists[m.group(1)] = [[x.strip().strip("'").strip() for x in n.split(",")] for n in names]
    par = open(os.path.join(REF, "parameters.F90")).read()
    attrs = re.findall(r"call nc_write_att_(\w+)\(.*?,\s*'([^']*)'\)", par)
    out = {"groups": groups, "dimensions": dims, "variables": variables, "snapshot_variables": lists,
           "global_attributes": sorted({(n, t) for t, n in attrs})}
    json.dump(out, open(os.path.join(HERE, "nc_schema_reference.json"), "w"), indent=1)
    print(len(groups), "groups", len(dims), "dimensions", len(variables), "variables", len(out["global_attributes"]), "attributes")


if __name__ == "__main__":
    main()
