"""Extracts the known-answer vectors of SURVEY.md Appendix F (printed by the reference's own Som.cpp /
Transformation.cpp during the survey, %.9g => round-trips f32) into tests/golden/appendix_f.json.

Usage: python tests/golden/make_appendix_f.py
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def floats(s):
    return [float(t) for t in s.replace(",", " ").split()]


def main():
    text = open(os.path.join(REPO, "SURVEY.md")).read()
    app = text[text.index("## Appendix F"):]
    blocks = []
    cur = None
    for line in app.splitlines():
        m = re.match(r"== (\w+)/(\w+), W=(\d+) H=(\d+) Din=(\d+) Dm=(\d+), randomInitialize\((\d+),([\d.]+)f\)", line)
        if m:
            cur = dict(transform=m.group(1), decay=m.group(2), W=int(m.group(3)), H=int(m.group(4)), Din=int(m.group(5)), Dm=int(m.group(6)),
                       seed=int(m.group(7)), init_sigma=float(m.group(8)), initial={}, steps=[], final={}, U=None)
            blocks.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+(\d+): \[(.*)\]\s*$", line)
        if m:
            cur["initial"][m.group(1)] = floats(m.group(2))
            continue
        m = re.match(r"\s+step (\d+): v=\[(.*?)\] eta=([\d.]+) sigma=([\d.]+) -> bmu=\((\d+),(\d+)\) lastBMU=(\d+) distErr=(\S+) resid2=(\S+)", line)
        if m:
            cur["steps"].append(dict(v=floats(m.group(2)), eta=float(m.group(3)), sigma=float(m.group(4)), bmu_x=int(m.group(5)), bmu_y=int(m.group(6)),
                                     lastBMU=int(m.group(7)), distErr=float(m.group(8)), resid2=float(m.group(9))))
            continue
        m = re.match(r"\s+bmu node (\d+): M=\[(.*?)\] sigma=\[(.*?)\] W=(\S+)", line)
        if m:
            cur["steps"][-1]["bmu_node"] = dict(node=int(m.group(1)), M=floats(m.group(2)), sigma=floats(m.group(3)), W=float(m.group(4)))
            continue
        m = re.match(r"\s+final node (\d+): M=\[(.*?)\] sigma=\[(.*?)\] W=(\S+)", line)
        if m:
            cur["final"][m.group(1)] = dict(M=floats(m.group(2)), sigma=floats(m.group(3)), W=float(m.group(4)))
            continue
        m = re.match(r"\s+U: (.*)$", line)
        if m:
            cur["U"] = floats(m.group(1))
            continue
    out = dict(blocks=blocks)
    m = re.search(r"nw\(0,0\|1,2;sigma=1.5\)=(\S+) nw\(3,1\|1,2;2.0\)=(\S+) nw\(1,2\|1,2;1.0\)=(\S+) nw\(0,2\|1,2;1.0\)=(\S+)", app)
    out["nw"] = [dict(c=[0, 0], b=[1, 2], sigma=1.5, value=float(m.group(1))), dict(c=[3, 1], b=[1, 2], sigma=2.0, value=float(m.group(2))),
                 dict(c=[1, 2], b=[1, 2], sigma=1.0, value=float(m.group(3))), dict(c=[0, 2], b=[1, 2], sigma=1.0, value=float(m.group(4)))]
    sc = {}
    sc["find_bmu"] = [dict(row=int(a), x=int(b), y=int(c), dist=float(d), distRaw=float(e))
                      for a, b, c, d, e in re.findall(r"findBmu\(row (\d)\)=\((\d+),(\d+)\) dist=(\S+) distRaw=(\S+)", app)]
    sc["evaluate"] = float(re.search(r"\nevaluate=(\S+)", app).group(1))
    sc["bmuHits"] = [int(t) for t in re.search(r"bmuHits: ([\d ]+)", app).group(1).split()]
    sc["restricted"] = [dict(row=int(a), minHits=int(b), x=int(c), y=int(d))
                        for a, b, c, d in re.findall(r"findRestrictedBmu\(row (\d),minHits=(\d)\)=\((\d+),(\d+)\)", app)]
    sc["bmd_row0_minHits0"] = floats(re.search(r"findRestrictedBmd\(row0,minHits=0\): (.*)", app).group(1))
    m = re.search(r"measureSimilarity\(k=1,minHits=0\)=(\d)\s+\(k=3\)=(\d)", app)
    sc["measureSimilarity"] = dict(k1=int(m.group(1)), k3=int(m.group(2)))
    out["scoring_on_standard_exponential_final_map"] = sc
    with open(os.path.join(HERE, "appendix_f.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("blocks:", [(b["transform"], b["decay"], len(b["steps"]), len(b["final"])) for b in blocks])


if __name__ == "__main__":
    main()
