#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the reference's own test fixtures and known answers.

Run in the build container (needs /root/reference, which the GPU box does not have):
    python tests/golden/make_golden.py
Each fixture carries the character matrix (re-encoded, not a copy of the file), the Newick
string of the reference test and the expected values, with the reference file:line they
come from.  The numbers are the reference's published known answers; nothing here is
computed by this repo's own likelihood code.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/test"

from mcphylo_jl_b200.parser import ParseCSV, ParseNexus  # noqa: E402


def matrix_fixture(path):
    ntax, nchar, gap, miss, symbols, df, langs = ParseNexus(path)
    return {"ntax": ntax, "nchar": nchar, "gap": gap, "missing": miss, "symbols": symbols,
            "taxa": langs, "rows": ["".join(r) for r in df]}


def main():
    primates = matrix_fixture(f"{REF}/likelihood/primates.nex")
    primates.update({
        "source": "test/likelihood/primates.nex; test/likelihood/felsenstein.jl:4-20; "
                  "test/distributions/phylodist.jl:3-15,41,86-94",
        "newick": "(((Tarsius_syrichta:0.0510942,(Lemur_catta:0.0136013,Homo_sapiens:0.0370755)12:0.0343822)"
                  "13:0.224569,(Pan:0.0712342,Gorilla:0.03754)14:0.0295151)15:0.0768634,((Pongo:"
                  "0.020513,Hylobates:0.159117)16:0.239429,Macaca_fuscata:0.454752)"
                  "17:0.0902988,((M_mulatta:0.0644278,M_fascicularis:0.318016)"
                  "18:0.015879,(M_sylvanus:0.100663,Saimiri_sciureus:0.0112774)19:0.2727)20:0.0448203);",
        "model": "JC", "base_freq": [0.25] * 4, "substitution_rates": [1.0], "rates": [1.0],
        "logpdf": -8677.360274116634,            # felsenstein.jl:20
        "size": [4, 1, 22],                      # phylodist.jl:41
    })
    simudata = matrix_fixture(f"{REF}/likelihood/simudata.nex")
    simudata.update({
        "source": "test/likelihood/simudata.nex; test/likelihood/felsenstein.jl:23-41",
        "newick": "(((0:0.110833,1:0.0137979)10:0.146124,(2:0.197891,(3:0.132967,(4:0.0378759,5:0.089252)"
                  "11:0.101833)12:0.184301)13:0.0450774)14:0.335725,6:0.153197,(7:0.0216218,(8:0.0781687,"
                  "9:0.120419)15:0.0209114)16:0.0209771);",
        "model": "JC", "base_freq": [0.25] * 4, "substitution_rates": [1.0], "rates": [1.0],
        # felsenstein.jl:35 — stale by 1.95e-9 relative; passes only at the
        # reference's own isapprox tolerance sqrt(eps) ~ 1.5e-8
        "logpdf_loose": -738.7363911756175,
        "logpdf_rtol": 1.5e-8,
        # felsenstein.jl:37, order = node.num 1..17
        "grad": [-56.25542148325748, -38.05203887880975, 48.05792385792187, -13.52161136915132,
                 -6.157297983096069, -19.400861758279206, -13.271503408059996, 23.78407661010017,
                 -2.7933830668575474, -10.342020138166133, -6.786353636249929, -32.05726418827399,
                 -3.742827673853779, 5.710607498000256, -2.174093469729043, 106.0017224580945,
                 57.19513048979947],
    })
    # parser fixture: test/parser/example.csv + expectations of test/parser/parser.jl:4-10
    ntax, nchar, gap, miss, symbols, df, langs = ParseCSV(f"{REF}/parser/example.csv", "-", "?", True)
    parser = {
        "source": "test/parser/example.csv (header=true drops the first taxon); test/parser/parser.jl:4-10",
        "taxa": langs, "rows": ["".join(r) for r in df], "gap": gap, "missing": miss, "symbols": symbols,
        "slot1": [[1.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0],
                  [0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0]],
        "slot2": [[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1.0],
                  [0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 0.0]],
    }
    models = {
        "source": "test/likelihood/substitutionmodels.jl:3-14,28-36,50-77; test/likelihood/rates.jl:3-18",
        "Restriction": {"base_freq": [0.3, 0.7],
                        "U": [[0.7, 0.7], [-0.3, 0.7]], "D": [-1.0, 0.0],
                        "Uinv": [[1.0, -1.0], [0.428571428, 1.0]], "mu": 2.380952380952381},
        "JC3": {"base_freq": [0.15, 0.45, 0.4], "mu": 1.5,
                "Q": [[-2 / 3, 1 / 3, 1 / 3], [1 / 3, -2 / 3, 1 / 3], [1 / 3, 1 / 3, -2 / 3]]},
        "GTR": {"base_freq": [0.5, 0.2, 0.3],
                "rates": [0.060325906174435326, 0.48940696298364417, 2.4502671308419206],
                "U": [[-0.06919383208280398, 0.5545230380047614, 0.5773502691896261],
                      [-0.7729814547853464, -0.6912032799070211, 0.5773502691896257],
                      [0.6306440233282372, -0.46340287673658875, 0.5773502691896257]],
                "D": [-1.3622649268468585, -0.29662234328312154, -5.551115123125783e-17],
                "Uinv": [[-0.1434321592442829, -0.6409264859501984, 0.7843586451944811],
                         [0.8837782123919058, -0.44064564120807803, -0.443132571183828],
                         [0.8660254037844387, 0.34641016151377513, 0.5196152422706631]],
                "mu": 0.6028137161614641},
        "freeK": {"rates": [0.03338775337571049, 0.2519159175077897, 0.8202684796606095, 2.8944278494558904],
                  "U": [[-0.9913312346241385, -0.7071067811865475], [0.131386389167909, -0.7071067811865476]],
                  "D": [-0.2853036708835002, -6.938893903907228e-18],
                  "Uinv": [[-0.8906959139221835, 0.8906959139221834], [-0.16549879465230727, -1.2487147677207877]],
                  "mu": 3.5050372710007514},
        "setmatrix": {"in": [0.1, 0.2, 0.3, 0.4, 0.5, 0.6],
                      "out": [[0.0, 0.1, 0.2, 0.4], [0.1, 0.0, 0.3, 0.5], [0.2, 0.3, 0.0, 0.6], [0.4, 0.5, 0.6, 0.0]]},
        "gamma_mean": {"args": [0.5, 0.5, 4],
                       "out": [0.03338775337571049, 0.2519159175077897, 0.8202684796606095, 2.8944278494558904]},
        "gamma_median": {"args": [0.5, 0.5, 4],
                         "out": [0.029077754761925846, 0.2807145371399754, 0.92477306511421, 2.7654346429838887]},
        "median_boundaries": {"args": [0.5, 0.5, 4],
                              "out": [0.024746651492520606, 0.23890238006213632, 0.7870290171790321, 2.353525844604548]},
        "mean_boundaries": {"args": [0.5, 0.5, 4],
                            "out": [0.10153104426762159, 0.45493642311957283, 1.3233036969314669]},
    }
    for name, obj in (("primates", primates), ("simudata", simudata), ("parser_csv", parser),
                      ("models", models)):
        with open(os.path.join(HERE, f"{name}.json"), "w") as fh:
            json.dump(obj, fh, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
