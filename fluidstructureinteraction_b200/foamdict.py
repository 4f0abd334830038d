"""Reader for foam-extend ASCII dictionaries (the case inputs of the reference: 0/D, 0/pointD, constant/rheologyProperties,
constant/stressProperties, system/fvSolution, system/fvSchemes, system/controlDict, system/decomposeParDict ...), i.e. the
subset of [EXT] `dictionary`/`ITstream` syntax those files use: `key value ... ;`, `key { ... }`, lists `( ... )` and
`N ( ... )`, dimension sets `[1 -3 0 0 0 0 0]`, `//` and `/* */` comments, the FoamFile header, keys such as
`laplacian(DD,D)` or `interpolate(grad(U))`.  INPUT handling for tests / benchmarks, not part of the timed path.

parse(text) -> nested dict; an entry's value is a scalar (int / float / str) when it has one token, else the list of its
tokens; a parenthesised list becomes a Python list, a dimension set a tuple.  A key that occurs twice keeps the last value
(as foam-extend's dictionary does on merge)."""
import re

_TOKEN = re.compile(r"""
    "(?:[^"\\]|\\.)*"            |   # quoted string
    [{}()\[\];]                  |   # punctuation
    [^\s{}()\[\];"]+                 # word / number
""", re.X)


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def _atom(t):
    if t.startswith('"'):
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t)
    except ValueError:
        return t


def _join_keyword(tokens, i):
    """A keyword may contain balanced parentheses with no white space: laplacian(DD,D), interpolate(grad(U)).  The
    tokenizer split them; glue the pieces back while the parentheses directly follow the word."""
    word = tokens[i][0]
    j = i + 1
    while j < len(tokens) and tokens[j][0] == "(" and tokens[j][1] == tokens[j - 1][2]:     # '(' adjacent to the word
        depth = 0
        while j < len(tokens):
            t = tokens[j][0]
            word += t
            depth += (t == "(") - (t == ")")
            j += 1
            if depth == 0:
                break
    return word, j


def _parse_list(tokens, i, close):
    out = []
    while tokens[i][0] != close:
        t = tokens[i][0]
        if t == "(":
            v, i = _parse_list(tokens, i + 1, ")")
            out.append(v)
        elif t == "[":
            v, i = _parse_list(tokens, i + 1, "]")
            out.append(tuple(v))
        elif t == "{":
            v, i = _parse_dict(tokens, i + 1, "}")
            out.append(v)
        else:
            # `word (` with the parenthesis attached is a function-like token (hex (0 1 ...) is NOT: there is white space)
            if i + 1 < len(tokens) and tokens[i + 1][0] == "(" and tokens[i + 1][1] == tokens[i][2] and not re.match(r"^[-+0-9.]", t):
                w, i = _join_keyword(tokens, i)
                out.append(w)
            else:
                out.append(_atom(t)); i += 1
    return out, i + 1


def _parse_dict(tokens, i, close):
    d = {}
    while i < len(tokens) and tokens[i][0] != close:
        if tokens[i][0] == "\0":                 # end of file inside a sub-dictionary (run/fsiFoam/3dTube/fluid/0/p lacks the
            return d, i                          # closing brace of boundaryField): close what is open, as the file's reader must
        key, i = _join_keyword(tokens, i)
        if key == ";":
            continue
        if key.startswith('"'):
            key = key[1:-1]
        if tokens[i][0] == "{":
            d[key], i = _parse_dict(tokens, i + 1, "}")
            continue
        vals = []
        while tokens[i][0] != ";":
            t = tokens[i][0]
            if t == "(":
                v, i = _parse_list(tokens, i + 1, ")")
                vals.append(v)
            elif t == "[":
                v, i = _parse_list(tokens, i + 1, "]")
                vals.append(tuple(v))
            elif t == "{":                      # `key value { ... }` does not occur in the shipped files
                v, i = _parse_dict(tokens, i + 1, "}")
                vals.append(v)
            else:
                if i + 1 < len(tokens) and tokens[i + 1][0] == "(" and tokens[i + 1][1] == tokens[i][2] and not re.match(r"^[-+0-9.]", t):
                    w, i = _join_keyword(tokens, i)
                    vals.append(w)
                else:
                    vals.append(_atom(t)); i += 1
        i += 1
        d[key] = vals[0] if len(vals) == 1 else vals
    return d, i + 1


def parse(text, keep_header=False):
    text = _strip_comments(text)
    tokens = [(m.group(0), m.start(), m.end()) for m in _TOKEN.finditer(text)]
    d, _ = _parse_dict(tokens + [("\0", -1, -1)], 0, "\0")
    if not keep_header:
        d.pop("FoamFile", None)
    return d


def read(path, keep_header=False):
    with open(path, "r") as f:
        return parse(f.read(), keep_header)


def uniform(v, n, width):
    """A field entry `uniform (a b c)` / `uniform s` (tokens as parse() returns them) -> list of n rows."""
    if isinstance(v, list) and len(v) == 2 and v[0] == "uniform":
        x = v[1]
        row = [float(t) for t in x] if isinstance(x, list) else [float(x)]*width
        assert len(row) == width
        return [row[:] for _ in range(n)]
    raise ValueError("only `uniform` field entries are supported: %r" % (v,))


def dimensioned(v):
    """`rho rho [1 -3 0 0 0 0 0] 7854;` -> 7854.0 (name and dimensions dropped); a bare number passes through."""
    if isinstance(v, list):
        return float(v[-1])
    return float(v)


def switch(v):
    """[EXT] Switch: yes/no, on/off, true/false, y/n, t/f."""
    s = str(v).lower()
    if s in ("yes", "on", "true", "y", "t", "1"):
        return True
    if s in ("no", "off", "false", "n", "f", "0", "none"):
        return False
    raise ValueError("not a Switch: %r" % (v,))
