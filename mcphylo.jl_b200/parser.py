"""Input side of the path: NEXUS / CSV character matrices -> per-leaf data indexed by node.num.

Restates /root/reference/src/Parser/ParseNexus.jl:8-130, ParseCSV.jl:17-62 and
`datafortree` (/root/reference/src/Parser/Parser.jl:48-85).  Two targets:

  datafortree   the reference layout: Float64 (K, S, NN) column-major, leaf l's one-hot /
                all-ones columns at x[:, :, l.num-1]   (numpy: Fortran-ordered array)
  codesfortree  the compact device layout: uint8 (n_leaves, S) state codes, code K means
                gap/missing (all ones); row i belongs to leaf_nums[i]

Both use the same alphabet rule (sorted set of symbols unless the file declares them) and
the same error for unknown symbols.
"""
from __future__ import annotations

import warnings
from typing import List, Sequence, Tuple

import numpy as np

from .tree import GeneralNode, find_by_name, get_leaves, post_order


class FileSyntaxError(Exception):
    pass


def get_alphabet(df: np.ndarray, gap: str, missing_representation: str) -> List[str]:
    alphabet = set()
    for entry in np.asarray(df).ravel():
        entry = str(entry)
        if entry != gap and entry != missing_representation:
            alphabet.add(entry)
    return sorted(alphabet)


def extract_meta_info(content: List[str]):
    ntax = 0
    nchar = 0
    gap = "-"
    symbols = "NOSYMBOLS"
    missing_representation = "?"
    while True:
        if not content:
            raise FileSyntaxError("no matrix block found")
        line = content.pop(0)
        if "matrix" in line.lower():
            break
        for entry in line.split():
            info = [p.lower() for p in entry.split("=")]
            if len(info) != 1:
                choped = info[1]
                if choped.endswith(";"):
                    choped = choped[:-1]
                k_word = info[0]
                if k_word == "ntax":
                    ntax = int(choped)
                elif k_word == "nchar":
                    nchar = int(choped)
                elif k_word == "gap":
                    gap = choped
                elif k_word == "missing":
                    missing_representation = choped
                elif k_word == "symbols":
                    symbols = choped.strip('"')
                else:
                    warnings.warn(f"Keyword {k_word} not understood, will be ignored")
    return ntax, nchar, gap, missing_representation, symbols


def create_nexusdf(filecontent: List[str]) -> Tuple[List[str], np.ndarray]:
    languages: List[str] = []
    rows: List[str] = []
    while True:
        if not filecontent:
            raise FileSyntaxError("matrix block is not terminated by ';'")
        line = filecontent.pop(0)
        if line == "":
            continue
        if line[-1] == ";":
            break
        lang, raw = line.split(None, 1)
        raw = "".join(raw.strip())
        languages.append(lang)
        rows.append(raw)
    width = len(rows[-1]) if rows else 0
    if any(len(r) != width for r in rows):
        raise FileSyntaxError("rows of the character matrix differ in length")
    df = np.array([list(r) for r in rows], dtype="<U1").reshape(len(rows), width)
    return languages, df


def ParseNexus(filename: str):
    """Returns (ntax, nchar, gap, missing, symbols, df, langs) like the reference."""
    with open(filename, "r") as fh:
        content = fh.read().splitlines()
    if not content or content[0].lower() != "#nexus":
        raise FileSyntaxError(f"{filename} is not a Nexus file!")
    while True:
        if not content:
            raise FileSyntaxError(f"{filename} has no data block")
        line = content.pop(0)
        if line.lower() == "begin data;":
            break
    ntax, nchar, gap, miss, symbols = extract_meta_info(content)
    langs, df = create_nexusdf(content)
    out_symbols = get_alphabet(df, gap, miss) if symbols == "NOSYMBOLS" else [s for s in symbols]
    return ntax, nchar, gap, miss, out_symbols, df, langs


def create_csvdf(filecontent: Sequence[str], separator: str = ","):
    language: List[str] = []
    rows: List[List[str]] = []
    for line in filecontent:
        if line == "":
            continue
        parts = line.split(separator)
        language.append(parts[0])
        rows.append([p[0] for p in parts[1:]])
    df = np.array(rows, dtype="<U1")
    return language, df


def ParseCSV(filename: str, gap: str, miss: str, header: bool = True):
    with open(filename, "r") as fh:
        content = fh.read().splitlines()
    if header:
        content.pop(0)
    langs, df = create_csvdf(content)
    ntax, nchar = df.shape
    symbols = get_alphabet(df, gap, miss)
    return ntax, nchar, gap, miss, symbols, df, langs


def _code_matrix(df: np.ndarray, symbols: Sequence[str], gap: str, miss: str) -> np.ndarray:
    """(n_rows, S) uint8 state codes; K = gap/missing.  Unknown symbol -> error (Parser.jl:76)."""
    K = len(symbols)
    if K > 254:
        raise ValueError("more than 254 states cannot be held in uint8 codes")
    df = np.asarray(df)
    codes = np.full(df.shape, 255, dtype=np.uint8)
    for i, s in enumerate(symbols):
        # first match wins, like findfirst
        codes[(df == s) & (codes == 255)] = i
    codes[(codes == 255) & ((df == gap) | (df == miss))] = K
    bad = np.argwhere(codes == 255)
    if bad.size:
        r, c = bad[0]
        raise ValueError(f"unknown symbol {df[r, c]}, {list(symbols)}")
    return codes


def datafortree(df, leave_names: Sequence[str], tree: GeneralNode, symbols: Sequence[str],
                gap: str, miss: str, log_space: bool = False) -> np.ndarray:
    """Dense reference layout (K, S, NN), Fortran order; internal slots are zero."""
    n_nodes = len(post_order(tree))
    K = len(symbols)
    codes = _code_matrix(df, symbols, gap, miss)
    S = codes.shape[1]
    x = np.zeros((K, S, n_nodes), dtype=np.float64, order="F")
    one, zero = (0.0, -np.inf) if log_space else (1.0, 0.0)
    for row, name in enumerate(leave_names):
        num = find_by_name(tree, name).num
        slot = np.full((K, S), zero)
        c = codes[row]
        known = c < K
        slot[c[known], np.nonzero(known)[0]] = one
        slot[:, ~known] = one
        x[:, :, num - 1] = slot
    return x


def codesfortree(df, leave_names: Sequence[str], tree: GeneralNode, symbols: Sequence[str],
                 gap: str, miss: str) -> Tuple[np.ndarray, np.ndarray]:
    """Compact layout: (codes uint8 (n_leaves, S), leaf_nums int32 (n_leaves,)); rows follow
    get_leaves(tree) order so that row i is leaf_nums[i]."""
    codes = _code_matrix(df, symbols, gap, miss)
    by_name = {name: i for i, name in enumerate(leave_names)}
    leaves = get_leaves(tree)
    out = np.empty((len(leaves), codes.shape[1]), dtype=np.uint8)
    nums = np.empty(len(leaves), dtype=np.int32)
    for i, leaf in enumerate(leaves):
        out[i] = codes[by_name[leaf.name]]
        nums[i] = leaf.num
    return out, nums


def dense_to_codes(x: np.ndarray, leaf_nums: Sequence[int]) -> np.ndarray:
    """Classify every leaf column of a dense (K, S, NN) array: one-hot -> state, all-ones ->
    K.  Anything else raises; the device path stores indicator data only."""
    K, S, _ = x.shape
    out = np.empty((len(leaf_nums), S), dtype=np.uint8)
    for i, num in enumerate(leaf_nums):
        col = np.asarray(x[:, :, num - 1])
        ones = col == 1.0
        zeros = col == 0.0
        n1 = ones.sum(axis=0)
        ok_onehot = (n1 == 1) & (zeros.sum(axis=0) == K - 1)
        ok_gap = n1 == K
        if not np.all(ok_onehot | ok_gap):
            s = int(np.argmin(ok_onehot | ok_gap))
            raise ValueError(f"leaf {num}, site {s + 1}: column is neither one-hot nor all ones")
        out[i] = np.where(ok_gap, K, np.argmax(ones, axis=0)).astype(np.uint8)
    return out
