"""python tests/golden/make_pin_kit.py -- writes the cmdstanr pin kit: for every golden fixture (tests/golden/*.npz) the Stan
`data` list of the reference program as CmdStan JSON and the unconstrained vectors the fixture was evaluated at.

Why: the oracle's parity is "unpinned"  because neither R nor CmdStan exists in this image.  Somebody
with an R + cmdstanr installation runs

    Rscript scripts/pin_with_cmdstanr.R /path/to/bayesVG/inst tests/golden/pin

which evaluates `$log_prob()` / `$grad_log_prob()` of inst/approxGP{,2,3,4}.stan at exactly these vectors and drops
`<name>.cmdstanr.json` next to the inputs; tests/test_pin_cmdstanr.py then compares the oracle (CPU suite) and the CUDA engine
(GPU suite) with those numbers.  The data list follows R/findSpatiallyVariableFeaturesBayes.R:293-331 of the reference: long
format, gene-major (gene_id varies slowest), 1-based ids."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PROGRAM = {"gp": "approxGP.stan", "gp4": "approxGP4.stan", "gp2": "approxGP2.stan", "gp3": "approxGP3.stan"}


def main():
    out = os.path.join(HERE, "pin")
    os.makedirs(out, exist_ok=True)
    index = []
    for fn in sorted(os.listdir(HERE)):
        if not fn.endswith(".npz") or not fn.startswith("gp"):
            continue
        name = fn[:-4]
        d = np.load(os.path.join(HERE, fn))
        M, G, k = int(d["M"]), int(d["G"]), int(d["k"])
        y = np.asarray(d["y"])          # M x G
        nb = name.split("_")[0] in ("gp3", "gp4")
        data = {"M": M, "N": M * G, "G": G, "k": k,
                "spot_id": [s + 1 for _ in range(G) for s in range(M)],
                "gene_id": [g + 1 for g in range(G) for _ in range(M)],
                "phi": np.asarray(d["phi"]).tolist(),
                "y": ([int(v) for v in y.T.ravel()] if nb else [float(v) for v in y.T.ravel()])}
        if "gene_depths" in d.files:
            data["gene_depths"] = [float(v) for v in d["gene_depths"]]
        json.dump(data, open(os.path.join(out, f"{name}.data.json"), "w"))
        json.dump({"program": PROGRAM[name.split("_")[0]], "thetas": np.asarray(d["thetas"]).tolist()},
                  open(os.path.join(out, f"{name}.theta.json"), "w"))
        index.append({"name": name, "program": PROGRAM[name.split("_")[0]], "M": M, "G": G, "k": k, "n_theta": int(d["thetas"].shape[0])})
    json.dump(index, open(os.path.join(out, "index.json"), "w"), indent=1)
    print(f"wrote {len(index)} data / theta pairs to {out}")


if __name__ == "__main__":
    main()
