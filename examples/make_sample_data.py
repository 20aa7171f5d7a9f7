"""Write a small synthetic Criteo-shaped CSV (label, I1..I13, C1..C26) -- stands in for the
reference's examples/train100.csv (100 rows of the Kaggle sample, already label-encoded)."""
import argparse

import numpy as np
import pandas


def make(rows=100, seed=0, path=None):
    rng = np.random.default_rng(seed)
    cols = {"label": rng.integers(0, 2, rows)}
    for i in range(1, 14):
        cols["I%d" % i] = rng.random(rows).round(6)
    vocab = [1460, 583, 100000, 50000, 305, 24, 12517, 633, 3, 93145, 5683, 80000, 3194, 27, 14992, 60000, 10, 5652,
             2173, 4, 70000, 18, 15, 28618, 105, 14257]
    for i, v in enumerate(vocab, start=1):
        cols["C%d" % i] = rng.integers(0, v, rows)
    df = pandas.DataFrame(cols)
    if path:
        df.to_csv(path)
    return df


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100)
    ap.add_argument("--out", default="train100.csv")
    a = ap.parse_args()
    make(a.rows, path=a.out)
