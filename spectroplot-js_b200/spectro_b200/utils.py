"""lookup(): name resolution of windows / colormaps (reference lib/utils.js:25-40)."""


def lookup(table, array_or_key):
    """Exact key, else case-insensitive equality, else case-insensitive PREFIX match in key
    order; non-strings pass through; unknown names give None."""
    if not array_or_key:
        return array_or_key
    if not isinstance(array_or_key, str):
        return array_or_key
    if array_or_key in table and table[array_or_key]:
        return table[array_or_key]
    match = array_or_key.lower()
    for key in table:
        if key.lower() == match:
            return table[key]
    for key in table:
        if key.lower().startswith(match):
            return table[key]
    return None


def js_parse_int(value, default):
    """`parseInt(value, 10) || default` (lib/spectroplot.js:239-250)."""
    import re
    if isinstance(value, bool):
        return default
    if isinstance(value, (int, float)):
        if value != value or value in (float("inf"), float("-inf")):
            return default
        v = int(value)          # parseInt(String(number)) truncates
        return v or default
    m = re.match(r"\s*([+-]?\d+)", str(value))
    if not m:
        return default
    return int(m.group(1)) or default
