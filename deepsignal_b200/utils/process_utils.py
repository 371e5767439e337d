"""Host-side sequence helpers for the call_mods / extract path.

Behavioural mirror of the few helpers of the reference that sit on the hot
path (reference: deepsignal/utils/process_utils.py:21-22 base/code tables,
:52-54 str2bool, :95-110 motif scan, :113-143 IUPAC expansion).  Written from
the documented behaviour, not translated.
"""
from __future__ import annotations

import itertools

base2code_dna = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}
code2base_dna = {v: k for k, v in base2code_dna.items()}
base2code_rna = {"A": 0, "C": 1, "G": 2, "U": 3, "N": 4}
code2base_rna = {v: k for k, v in base2code_rna.items()}

_IUPAC_DNA = {
    "A": "A", "C": "C", "G": "G", "T": "T",
    "R": "AG", "M": "AC", "S": "CG", "Y": "CT", "K": "GT", "W": "AT",
    "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT",
}
_IUPAC_RNA = {k: v.replace("T", "U") for k, v in _IUPAC_DNA.items() if k != "T"}
_IUPAC_RNA["U"] = "U"


def str2bool(v) -> bool:
    """'yes/true/t/1' (any case) -> True, anything else False (reference :52-54)."""
    return str(v).lower() in ("yes", "true", "t", "1")


def display_args(args) -> None:
    print("# ===============================================")
    print("## parameters: ")
    for key, val in vars(args).items():
        if key != "func":
            print("{}:\n\t{}".format(key, val))
    print("# ===============================================")


def get_motif_seqs(motifs: str, is_dna: bool = True) -> list[str]:
    """Expand a comma-separated list of IUPAC motifs into explicit sequences.

    Same result set as reference process_utils.py:113-143 (order is the
    lexicographic product order per motif; callers use it as a set).
    """
    table = _IUPAC_DNA if is_dna else _IUPAC_RNA
    out: list[str] = []
    for motif in motifs.strip().split(","):
        motif = motif.strip().upper()
        choices = [table[ch] for ch in motif]  # KeyError on a non-IUPAC letter, as the reference
        out.extend("".join(p) for p in itertools.product(*choices))
    return out


def get_refloc_of_methysite_in_motif(seqstr: str, motifset, methyloc_in_motif: int = 0) -> list[int]:
    """0-based locations (in seqstr) of the targeted base of every motif hit.

    Mirrors reference process_utils.py:95-110: all motifs share one length, a
    window slides by one base, hit at i emits i + methyloc_in_motif.
    """
    motifset = set(motifset)
    mlen = len(next(iter(motifset)))
    return [i + methyloc_in_motif
            for i in range(0, len(seqstr) - mlen + 1)
            if seqstr[i:i + mlen] in motifset]


def get_contig2len(ref_path: str) -> dict[str, int]:
    """FASTA contig name (up to first blank) -> length (reference utils/ref_reader.py:7-13)."""
    lens: dict[str, int] = {}
    name = None
    with open(ref_path, "r") as fh:
        for line in fh:
            if line.startswith(">"):
                name = line.strip()[1:].split(" ")[0]
                lens.setdefault(name, 0)
            elif name is not None:
                lens[name] += len(line.strip())
    return lens
