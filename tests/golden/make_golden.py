"""Extract the golden numbers printed in the reference's committed logs into a JSON fixture.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The fixture tests/golden/reference_logs.json is committed; the GPU box never reads
/root/reference.  Every entry records the log file:line it was taken from.
"""
import ast
import json
import os
import re

REF = "/root/reference/test"
HERE = os.path.dirname(os.path.abspath(__file__))


def lines(name):
    with open(os.path.join(REF, name)) as f:
        return f.read().split("\n")


def parse_array_block(ls, start):
    """Parse a numpy-printed array starting at line index ``start`` (0-based)."""
    depth = 0
    buf = []
    k = start
    while True:
        s = ls[k]
        buf.append(s)
        depth += s.count("[") - s.count("]")
        k += 1
        if depth == 0:
            break
    txt = " ".join(buf)
    txt = re.sub(r"\s+", " ", txt)
    txt = re.sub(r"(?<=[0-9.]) (?=[-0-9])", ", ", txt)
    txt = re.sub(r"\] \[", "], [", txt)
    return ast.literal_eval(txt.strip()), k


def main():
    out = {}
    tv = lines("testlog_vqe")
    out["ring_rhf"] = {"value": float(tv[1].split("=")[1]), "src": "test/testlog_vqe:2"}
    out["ring_fragment_spectrum_H_plus_0p75_S2"] = {
        "values": [float(tv[k].split(":")[1].split("au")[0]) for k in range(13, 19)],
        "s2": [float(tv[k].split("<S2>:")[1]) for k in range(13, 19)],
        "src": "test/testlog_vqe:14-19"}
    out["ring_hf_energy"] = {"value": float(tv[24].split(":")[1]), "src": "test/testlog_vqe:25"}
    wfn = {}
    for k in range(55, 91):
        m = re.match(r"\s*\((\d+), 0\)\s+\((.*)\)", tv[k])
        wfn[int(m.group(1))] = complex(m.group(2).replace("+0j", "")).real
    out["ring_vqe_wfn"] = {"index": list(wfn.keys()), "amp": list(wfn.values()), "src": "test/testlog_vqe:56-91"}
    out["ring_strings"] = {"values": [int(v) for v in tv[91].strip("[] ").split()], "src": "test/testlog_vqe:92"}
    fc = ast.literal_eval(tv[93])
    out["ring_fcivec"] = {"values": [complex(v).real for v in fc], "src": "test/testlog_vqe:94"}
    out["ring_vqe_imp_energy_float32_rdm"] = {"value": float(tv[96].split(":")[1]), "src": "test/testlog_vqe:97"}
    rdm1, k = parse_array_block(tv, 98)
    out["ring_vqe_rdm1"] = {"values": rdm1, "src": "test/testlog_vqe:99-102"}
    assert tv[k].startswith("2-rdm")
    rdm2, k = parse_array_block(tv, k + 1)
    out["ring_vqe_rdm2"] = {"values": rdm2, "src": "test/testlog_vqe:104-185"}

    tf = lines("testlog_fci")
    out["ring_fci"] = {
        "e_fci": [float(tf[k].split("=")[1]) for k in (7, 111, 215, 319)],
        "e_imp": [float(tf[k]) for k in (99, 203, 307, 411)],
        "e_tot": [float(tf[k]) for k in (105, 209, 313, 417)],
        "mu_nelec": [list(ast.literal_eval(tf[k].split("=")[1].strip())) for k in (106, 210, 314, 418)],
        "dmet_energy": float(tf[437].split(":")[1]),
        "src": "test/testlog_fci:8,100,106-107,112,204,...,438"}

    hv = lines("log_hchain_vqe")
    out["chain06_rhf"] = {"value": float(hv[1].split("=")[1]), "src": "test/log_hchain_vqe:2"}
    out["chain06_frag0_spectrum_H_plus_0p75_S2"] = {
        "values": [float(hv[k].split(":")[1].split("au")[0]) for k in range(13, 19)],
        "src": "test/log_hchain_vqe:14-19"}

    hf = lines("log_hchain_fci")
    out["chain_rhf_sweep"] = {"values": ast.literal_eval(hf[-3].split(":", 1)[1].strip()),
                              "src": "test/log_hchain_fci (HF energy list)", "bond": "0.6..3.0 step 0.1"}
    hd = lines("log_hchain_dmet_fci")
    out["chain_dmet_fci_sweep"] = {"values": ast.literal_eval(hd[54949].split(":", 1)[1].strip()),
                                   "src": "test/log_hchain_dmet_fci:54950"}
    out["chain06_dmet_fci_first"] = {
        "e_fci_frag0": float(hd[13].split("=")[1]),
        "src": "test/log_hchain_dmet_fci:14"}
    av = lines("atestlog_vqe")
    out["ring_dmet_vqe_current_code"] = {"value": float(av[345].split(":")[1]), "src": "test/atestlog_vqe:346"}

    with open(os.path.join(HERE, "reference_logs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
