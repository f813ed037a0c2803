"""Resolve a set of compile-time A/B switches in a source file (a tiny unifdef): `#if` / `#elif` / `#else` /
`#endif` groups whose condition only involves the given macros are evaluated and the dead branches removed, and
the `#ifndef M / #define M v / #endif` default blocks of those macros are dropped.
usage: python scripts/unifdef.py file.cu M1=v1 M2=v2 ..."""
import re, sys
path = sys.argv[1]
defs = dict(a.split("=") for a in sys.argv[2:])
lines = open(path).read().split("\n")
out, stack = [], []          # stack entries: [known, taken_any, active_now, parent_emit]
def emit(): return all(s[2] for s in stack if s[0]) if stack else True
def try_eval(expr):
    expr = re.sub(r"/\*.*", "", expr).strip()
    names = set(re.findall(r"[A-Za-z_]\w*", expr)) - {"defined"}
    if not names or not names <= set(defs): return None
    e = expr
    for n in names: e = re.sub(r"\b%s\b" % n, defs[n], e)
    e = e.replace("&&", " and ").replace("||", " or ").replace("!", " not ").replace(" not =", "!=")
    return bool(eval(e))
i = 0
while i < len(lines):
    ln = lines[i]; st = ln.strip()
    m = re.match(r"#\s*ifndef\s+(\w+)", st)
    if m and m.group(1) in defs and emit():            # default-definition block: drop it (with a continued comment)
        j = i + 1
        while not lines[j].strip().startswith("#endif"): j += 1
        i = j + 1; continue
    if re.match(r"#\s*if(n?def)?\b", st):
        v = try_eval(st.split(None, 1)[1]) if re.match(r"#\s*if\b", st) else None
        if v is None: stack.append([False, False, True]);  out.append(ln) if emit() else None
        else: stack.append([True, v, v])
        i += 1; continue
    if re.match(r"#\s*elif\b", st) and stack:
        s = stack[-1]
        if s[0]:
            v = try_eval(st.split(None, 1)[1]); s[2] = (not s[1]) and bool(v); s[1] = s[1] or s[2]
        elif emit(): out.append(ln)
        i += 1; continue
    if re.match(r"#\s*else\b", st) and stack:
        s = stack[-1]
        if s[0]: s[2] = not s[1]; s[1] = True
        elif emit(): out.append(ln)
        i += 1; continue
    if re.match(r"#\s*endif\b", st) and stack:
        s = stack.pop()
        if not s[0] and emit(): out.append(ln)
        i += 1; continue
    if emit(): out.append(ln)
    i += 1
open(path, "w").write("\n".join(out))
print(path, len(lines), "->", len(out), "lines")
