"""Re-derives every extracted force-field table in lightdock-rust_b200/data/ from the reference's Rust sources with a
parser that shares nothing with tools/extract_forcefield_tables.py (that tool is regex based; this one is a small
hand-written tokenizer), so an extraction bug cannot hide behind the fact that the oracle and the product read the
same TSV files.  CPU only; runs where /root/reference exists (the build container) and is
skipped on the GPU box, where the reference tree is absent by contract.

    DFIRE typing      src/dfire.rs:18-46 (r3_to_numerical), :56-78 (ATOMNUMBER), :79-101 (ATOMRES), :49-53 (DIST_TO_BINS)
    DNA / pyDock      src/dna.rs:64-233, src/pydock.rs:66-237
"""
import os

import pytest

REF = "/root/reference/src"
DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lightdock-rust_b200", "data")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")


# ---- a tiny Rust-literal tokenizer: string literals, numbers, identifiers, punctuation -----------------------
def tokens(text):
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c.isspace():
            i += 1
        elif text.startswith("//", i):
            while i < n and text[i] != "\n":
                i += 1
        elif c == '"':
            j = i + 1
            while text[j] != '"':
                assert text[j] != "\\", "escape sequences are not expected in these tables"
                j += 1
            yield ("str", text[i + 1:j])
            i = j + 1
        elif c.isdigit() or (c in "+-" and i + 1 < n and (text[i + 1].isdigit() or text[i + 1] == ".")):
            j = i + 1
            while j < n and (text[j].isdigit() or text[j] in ".eE_" or (text[j] in "+-" and text[j - 1] in "eE")):
                j += 1
            yield ("num", text[i:j].replace("_", ""))
            i = j
        elif c.isalpha() or c == "_":
            j = i + 1
            while j < n and (text[j].isalnum() or text[j] == "_"):
                j += 1
            yield ("id", text[i:j])
            i = j
        elif text.startswith("=>", i):
            yield ("op", "=>")
            i += 2
        else:
            yield ("op", c)
            i += 1


def after(tok, *ids):
    """Index just past the first occurrence of the identifier sequence `ids`."""
    for k in range(len(tok) - len(ids)):
        if all(tok[k + d] == ("id", ids[d]) for d in range(len(ids))):
            return k + len(ids)
    raise AssertionError(f"{ids} not found")


def hashmap_entries(tok, name):
    """`static ref NAME: ... = hashmap![ "k" => v, ... ];` -> dict, later keys overriding earlier ones (HashMap::insert)."""
    k = after(tok, "static", "ref", name)
    while tok[k] != ("id", "hashmap"):
        k += 1
    assert tok[k + 1] == ("op", "!") and tok[k + 2] == ("op", "[")
    k += 3
    out = {}
    while tok[k] != ("op", "]"):
        kind, key = tok[k]
        assert kind == "str" and tok[k + 1] == ("op", "=>"), (name, tok[k:k + 3])
        out[key] = tok[k + 2][1]
        k += 3
        if tok[k] == ("op", ","):
            k += 1
    return out


def read_tsv(name, cols=2):
    out = {}
    with open(os.path.join(DATA, name)) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            p = line.rstrip("\n").split("\t")
            assert len(p) == cols, (name, line)
            out[tuple(p[:-1]) if cols > 2 else p[0]] = p[-1]
    return out


# ---- DFIRE -----------------------------------------------------------------------------------------------
def dfire_tables():
    tok = list(tokens(open(os.path.join(REF, "dfire.rs")).read()))
    # r3_to_numerical: match arms `"ALA" => 0,` up to the `_ =>` arm
    k = after(tok, "fn", "r3_to_numerical")
    r3 = {}
    while tok[k] != ("id", "_"):
        if tok[k][0] == "str" and tok[k + 1] == ("op", "=>"):
            r3[tok[k][1]] = int(tok[k + 2][1])
        k += 1
    atomnumber = {key: int(v) for key, v in hashmap_entries(tok, "ATOMNUMBER").items()}
    # ATOMRES: vec![ vec![...], ... ]
    k = after(tok, "static", "ref", "ATOMRES")
    while tok[k] != ("id", "vec"):
        k += 1
    k += 3  # vec ! [
    rows = []
    while tok[k] == ("id", "vec"):
        k += 3
        row = []
        while tok[k] != ("op", "]"):
            if tok[k][0] == "num":
                row.append(int(tok[k][1]))
            k += 1
        rows.append(row)
        k += 1
        if tok[k] == ("op", ","):
            k += 1
    k = after(tok, "const", "DIST_TO_BINS")
    while tok[k] != ("op", "="):
        k += 1
    k += 1
    while tok[k] != ("op", "["):
        k += 1
    bins = []
    k += 1
    while tok[k] != ("op", "]"):
        if tok[k][0] == "num":
            bins.append(int(tok[k][1]))
        k += 1
    return r3, atomnumber, rows, bins


def test_dfire_atom_types_tsv_matches_the_reference_source():
    r3, atomnumber, atomres, _ = dfire_tables()
    assert len(r3) == 22 and len(atomres) == 22 and all(len(r) == 14 for r in atomres)
    derived = {}
    for key, anum in atomnumber.items():
        # format!("{}{}", res_name, atom_name) (src/dfire.rs:135): the residue is the longest r3 key the string starts with
        res = max((r for r in r3 if key.startswith(r)), key=len)
        derived[(res, key[len(res):])] = atomres[r3[res]][anum]
    tsv = {k: int(v) for k, v in read_tsv("dfire_atom_types.tsv", 3).items()}
    assert len(tsv) == len(atomnumber)
    assert derived == tsv
    assert tsv[("MMB", "BJ")] == 167 and tsv[("ALA", "N")] == 74      # spot values named in SURVEY.md a5
    assert max(tsv.values()) == 167 and min(tsv.values()) == 0


def test_dfire_dist_to_bins_tsv_matches_the_reference_source():
    *_, bins = dfire_tables()
    tsv = read_tsv("dfire_dist_to_bins.tsv")
    assert [int(tsv[str(i)]) for i in range(len(tsv))] == bins
    assert len(bins) == 51 and bins[:4] == [1, 1, 1, 2] and bins[29] == 21


# ---- DNA / pyDock ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("table,tsv_name,numeric", [
    ("VDW_CHARGES", "vdw_energy.tsv", True), ("VDW_RADII", "vdw_radius.tsv", True),
    ("AMBER_TYPES", "amber_types.tsv", False), ("ELE_CHARGES", "ele_charges.tsv", True),
    ("NT_ELE_CHARGES", "nt_ele_charges.tsv", True)])
def test_dna_and_pydock_tables_match_the_reference_source(table, tsv_name, numeric):
    dna = hashmap_entries(list(tokens(open(os.path.join(REF, "dna.rs")).read())), table)
    pyd = hashmap_entries(list(tokens(open(os.path.join(REF, "pydock.rs")).read())), table)
    tsv = read_tsv(tsv_name)
    extra_name = tsv_name.replace(".tsv", "_pydock_extra.tsv")
    extra = read_tsv(extra_name) if os.path.exists(os.path.join(DATA, extra_name)) else {}
    conv = float if numeric else str
    assert {k: conv(v) for k, v in tsv.items()} == {k: conv(v) for k, v in dna.items()}
    # pyDock = the DNA table plus its own extra rows (src/pydock.rs:147-148,209-210)
    merged = dict(tsv)
    merged.update(extra)
    assert {k: conv(v) for k, v in merged.items()} == {k: conv(v) for k, v in pyd.items()}
