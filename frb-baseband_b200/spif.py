"""Bit-extraction recipes of the corner turn (host side).

In the reference the raw multi-BBC recording is split into one 2-channel VDIF file per IF by
jive5ab's `spif2file`, driven by /root/reference/spif2file.sh: a mode string
`VDIF_<payload>-<Mbps>-<nbbc>-<nbits>` (/root/reference/base2fil.sh:308-318) selects a recipe
`W>[b0,b1,b2,b3][...]...:0-(n-1)` (/root/reference/spif2file.sh:31-98) -- W-bit input words, and per
output IF the 4 input bits that become its (pol A lsb, pol A msb, pol B lsb, pol B msb).  With
`flipIF` neighbouring IFs swap recipes (/root/reference/spif2file.sh:117-131).  libb2f does that
corner turn on the GPU (b2f_params.raw_word_bits / raw_bits); this module only turns the
reference's notation into those parameters.
"""
from __future__ import annotations

import re

#: recorder layouts, keyed like spif2file.sh's mode strings without the payload/rate prefix:
#: (number of BBC channels, bits per sample) -> recipe.  One line per distinct wiring in
#: /root/reference/spif2file.sh:31-77 (VDIF, 2-bit modes).
_RECIPES = {
    (32, 2): "64>[16,17,48,49][0,1,32,33][18,19,50,51][2,3,34,35][20,21,52,53][4,5,36,37][22,23,54,55][6,7,38,39]"
             "[24,25,56,57][8,9,40,41][26,27,58,59][10,11,42,43][28,29,60,61][12,13,44,45][30,31,62,63][14,15,46,47]:0-15",
    (16, 2): "32>[16,17,24,25][0,1,8,9][18,19,26,27][2,3,10,11][20,21,28,29][4,5,12,13][22,23,30,31][6,7,14,15]:0-7",
    (8, 2): "16>[8,9,12,13][0,1,4,5][10,11,14,15][2,3,6,7]:0-3",
    (16, 1): "16>[8,12][0,4][9,13][1,5][10,14][2,6][11,15][3,7]:0-7",        # 1-bit samples: (pol A, pol B) per IF, :58-61
}
#: the 1024 Mbps 16-BBC modes use the mirrored pairing (spif2file.sh:44-52)
_RECIPE_16_1024 = "32>[24,25,16,17][8,9,0,1][26,27,18,19][10,11,2,3][28,29,20,21][12,13,4,5][30,31,22,23][14,15,6,7]:0-7"


def mode_string(payload: int, datarate_mbps: int, nbbc: int, nbits: int) -> str:
    """base2fil.sh:308-318"""
    return f"VDIF_{payload}-{datarate_mbps}-{nbbc}-{nbits}"


def parse_recipe(recipe: str):
    """'32>[16,17,24,25][0,1,8,9]...:0-7' -> (32, [[16,17,24,25],[0,1,8,9],...]).

    Mark5B recordings store sign and magnitude the other way round, so their recipes (spif2file.sh:79-93) start with
    the step `swap_sign_mag+`: the two bits of every 2-bit field change places before the extraction.  Extracting bit b
    of the swapped word is extracting bit b ^ 1 of the recorded word, which is what is returned for such a recipe."""
    recipe = recipe.strip()
    swap = recipe.startswith("swap_sign_mag+")
    if swap:
        recipe = recipe[len("swap_sign_mag+"):]
    m = re.match(r"\s*(\d+)>((?:\[[0-9,]+\])+):", recipe)
    if not m:
        raise ValueError(f"not a spif2file recipe: {recipe!r}")
    groups = [[int(x) for x in g.split(",")] for g in re.findall(r"\[([0-9,]+)\]", m.group(2))]
    if any(len(g) != len(groups[0]) for g in groups) or len(groups[0]) not in (2, 4):
        raise ValueError("only dual-polarisation recipes with 1-bit (2 bits per IF) or 2-bit samples (4 bits per IF) are supported")
    if swap:
        groups = [[b ^ 1 for b in g] for g in groups]
    return int(m.group(1)), groups


def frame_geometry(mode: str):
    """(frame_bytes, header_bytes, raw_format) of a mode string: VDIF_<payload>-... has 32-byte headers,
    MARK5B-... 16-byte headers and 10000-byte payloads (spif2file.sh:99-113)."""
    if mode.startswith("MARK5B"):
        return 10016, 16, 1
    m = re.match(r"VDIF_(\d+)-", mode)
    if not m:
        raise ValueError(f"Cannot determine frame sizes from {mode}")
    return int(m.group(1)) + 32, 32, 0


def recipe_for_mode(mode: str, nif: int, flip_if: bool = False):
    """(word_bits, bits per IF 1..nif) for a base2fil mode string; flip_if swaps neighbouring IFs.  Two bits per IF
    mean 1-bit samples (in_nbit = 1), four mean 2-bit samples."""
    m5 = re.match(r"MARK5B-(\d+)-(\d+)-(\d+)$", mode)
    if m5:                                                          # spif2file.sh:79-93
        rate, nbbc, nbits = (int(x) for x in m5.groups())
        if (rate, nbbc, nbits) not in ((1024, 16, 2), (1024, 8, 2), (2048, 16, 2), (2048, 32, 2)):
            raise ValueError(f"mode {mode} not implemented")
        W, groups = parse_recipe("swap_sign_mag+" + _RECIPES[(nbbc, nbits)])
        return W, _flip(groups[:nif], flip_if)
    m = re.match(r"VDIF_(\d+)-(\d+)-(\d+)-(\d+)$", mode)
    if not m:
        raise ValueError(f"mode {mode} not implemented")
    payload, rate, nbbc, nbits = (int(x) for x in m.groups())
    if (nbbc, nbits) == (16, 2) and rate == 1024:
        rec = _RECIPE_16_1024
    elif (nbbc, nbits) in _RECIPES:
        rec = _RECIPES[(nbbc, nbits)]
    else:
        raise ValueError(f"mode {mode} not implemented")
    W, groups = parse_recipe(rec)
    return W, _flip(groups[:nif], flip_if)


def _flip(groups, flip_if: bool):
    """spif2file.sh:117-131: with flipIF neighbouring IFs swap recipes"""
    if flip_if:
        groups = [groups[i + 1] if i % 2 == 0 and i + 1 < len(groups) else groups[i - 1] if i % 2 else groups[i]
                  for i in range(len(groups))]
    return groups
