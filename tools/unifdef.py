#!/usr/bin/env python
"""Remove the preprocessor branches of macros that are never defined (measured-and-rejected experiment flags).
usage: unifdef.py file MACRO [MACRO ...]   - rewrites `file` in place.  Handles `#ifdef M`, `#ifndef M` and `#if` lines
whose condition is an ||-combination of defined(M) terms over the given macros only; everything else is left as it is."""
import re, sys

def main():
    path, macros = sys.argv[1], set(sys.argv[2:])
    lines = open(path).read().split("\n")
    out = []
    # stack entries: (kind, keep_now) kind: 'ours' or 'other'; for ours: state in {'if_true','if_false'} and whether else seen
    stack = []
    def emitting():
        return all(e["emit"] for e in stack)
    for ln in lines:
        t = ln.strip()
        m_ifdef = re.match(r"#\s*ifdef\s+(\w+)\s*$", t)
        m_ifndef = re.match(r"#\s*ifndef\s+(\w+)\s*$", t)
        m_if = re.match(r"#\s*if\s+(.*)$", t)
        if m_ifdef or m_ifndef or (m_if and not t.startswith("#ifdef") and not t.startswith("#ifndef")):
            val = None
            if m_ifdef and m_ifdef.group(1) in macros: val = False
            elif m_ifndef and m_ifndef.group(1) in macros: val = True
            elif m_if:
                cond = m_if.group(1)
                terms = [x.strip() for x in cond.split("||")]
                ok = all(re.fullmatch(r"defined\s*\(\s*(\w+)\s*\)", x) and re.fullmatch(r"defined\s*\(\s*(\w+)\s*\)", x).group(1) in macros for x in terms)
                if ok: val = False
            if val is None:
                if emitting(): out.append(ln)
                stack.append({"ours": False, "emit": True})
            else:
                stack.append({"ours": True, "emit": val, "cond": val})
            continue
        if re.match(r"#\s*else\b", t):
            e = stack[-1]
            if e["ours"]: e["emit"] = not e["cond"]
            elif emitting(): out.append(ln)
            continue
        if re.match(r"#\s*elif\b", t):
            e = stack[-1]
            assert not e["ours"], "elif on a removed conditional is not supported: " + ln
            if emitting(): out.append(ln)
            continue
        if re.match(r"#\s*endif\b", t):
            e = stack.pop()
            if not e["ours"] and emitting(): out.append(ln)
            continue
        if emitting(): out.append(ln)
    assert not stack
    open(path, "w").write("\n".join(out))

if __name__ == "__main__":
    main()
