"""File-name metadata (reference lib/parseFreqRate.js:16-70)."""
import re

_FLOAT = re.compile(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)")


def _parse_float(s):
    m = _FLOAT.match(s)
    return float(m.group(0)) if m else float("nan")


def parseFreqRate(name=""):
    """`..._433.92M_250k.cu8` -> {"freq": 433920000.0, "rate": 250000.0}"""
    if not name or not isinstance(name, str):
        return {"freq": 0, "rate": 0}
    pos = name.rfind("/")
    if pos != -1:
        name = name[pos + 1:]
    freq, rate = 0, 1
    p = 0
    while p < len(name) - 1:
        if name[p] in "_- .":
            p += 1
            f = _parse_float(name[p:])
            if f != f:
                p += 1          # `continue` still runs the loop's ++p
                continue
            while p < len(name) and (("0" <= name[p] <= "9") or name[p] == "."):
                p += 1
            if p < len(name) and name[p] in "Mm":
                freq = f * 1000000.0
            if p < len(name) and name[p] in "kK":
                rate = f * 1000.0
        p += 1
    return {"freq": freq, "rate": rate}


def parseFormat(name=""):
    """Upper-cased file extension, '?' when there is none."""
    if not name or not isinstance(name, str):
        return "?"
    pos = name.rfind(".")
    if pos != -1:
        return name[pos + 1:].upper()
    return "?"
