#!/usr/bin/env python3
"""Extract the force-field PARAMETER DATA used by the scoring path into neutral TSV files.

The AMBER94 atom-type/charge/vdW tables (src/dna.rs:64-233, src/pydock.rs:147-148,209-210) and the
DFIRE residue/atom -> atom-type numbering (src/dfire.rs:18-46,56-101) are data, not code.  They are
read programmatically from the reference and written as one-entry-per-line TSV under
lightdock-rust_b200/data/, which is what both the product host layer and the oracle's setup read.
Later `"KEY" => v` entries override earlier ones, as `HashMap::insert` does in the reference macro.

Run once in the build container:   python tools/extract_forcefield_tables.py
"""
import os, re, sys

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "lightdock-rust_b200", "data")


def blocks(text):
    """name -> body text of every `static ref NAME ... = hashmap![ ... ];`"""
    out = {}
    for m in re.finditer(r"static ref (\w+)\s*:[^=]*=\s*hashmap!\[(.*?)\];", text, re.S):
        out[m.group(1)] = m.group(2)
    return out


def entries(body):
    d = {}
    for m in re.finditer(r'"([^"]+)"\s*=>\s*("([^"]*)"|[-+0-9.eE]+)', body):
        d[m.group(1)] = m.group(3) if m.group(3) is not None else m.group(2)
    return d


def write(name, d, header):
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        f.write("# " + header + "\n")
        for k, v in d.items():
            f.write(f"{k}\t{v}\n")
    print(f"{name}: {len(d)} rows")


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; tables are already committed")
    os.makedirs(OUT, exist_ok=True)
    dna = blocks(open(f"{REF}/dna.rs").read())
    pyd = blocks(open(f"{REF}/pydock.rs").read())
    for nm, fn in (("VDW_CHARGES", "vdw_energy.tsv"), ("VDW_RADII", "vdw_radius.tsv"),
                   ("AMBER_TYPES", "amber_types.tsv"), ("ELE_CHARGES", "ele_charges.tsv"),
                   ("NT_ELE_CHARGES", "nt_ele_charges.tsv")):
        a = entries(dna[nm])
        b = entries(pyd[nm])
        extra = {k: v for k, v in b.items() if k not in a}
        changed = {k for k in a if k in b and a[k] != b[k]}
        missing = {k for k in a if k not in b}
        assert not changed and not missing, (nm, changed, missing)
        write(fn, a, f"{nm} (src/dna.rs); key<TAB>value")
        if extra:
            write(fn.replace(".tsv", "_pydock_extra.tsv"), extra,
                  f"{nm} rows present only in src/pydock.rs; key<TAB>value")
    # DFIRE typing: (residue, atom) -> type
    src = open(f"{REF}/dfire.rs").read()
    r3 = dict(re.findall(r'"(\w+)"\s*=>\s*(\d+),', src[src.index("fn r3_to_numerical"):src.index("DIST_TO_BINS")]))
    atomnumber = entries(blocks(src)["ATOMNUMBER"])
    m = re.search(r"static ref ATOMRES[^=]*=\s*vec!\[(.*?)\];\s*\n\}", src, re.S)
    atomres = [[int(x) for x in row.split(",")] for row in re.findall(r"vec!\[([0-9, ]+)\]", m.group(1))]
    assert len(atomres) == 22 and all(len(r) == 14 for r in atomres)
    rows = {}
    for key, anum in atomnumber.items():
        # key is residue name (3 chars) + atom name
        res, atom = key[:3], key[3:]
        rows[f"{res}\t{atom}"] = atomres[int(r3[res])][int(anum)]
    write("dfire_atom_types.tsv", rows, "DFIRE type = ATOMRES[r3_to_numerical(res)][ATOMNUMBER[res+atom]] "
          "(src/dfire.rs:18-101); residue<TAB>atom<TAB>type")
    bins = re.search(r"DIST_TO_BINS: &\[usize\] = &\[(.*?)\];", src, re.S).group(1)
    bins = [int(x) for x in bins.replace("\n", " ").split(",") if x.strip()]
    write("dfire_dist_to_bins.tsv", {str(i): b for i, b in enumerate(bins)},
          "DIST_TO_BINS (src/dfire.rs:49-53); index<TAB>bin(1-based)")


if __name__ == "__main__":
    main()
