"""Extract the SHAPE of every decoder call site in the reference's Python callers (api/, cli/, web/):
file, line, callee, positional-argument kinds, keyword names -- no source text.  Run in the build
container (needs /root/reference):

    python tests/golden/extract_callsites.py

Writes tests/golden/ref_callsites.json, which tests/test_reference_callers.py replays against
llm_decoder (on CPU: argument binding; on the GPU: the real calls on a small model)."""
import ast
import json
import os
import sys

REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_callsites.json")
CTORS = {"CUDADecoder", "INT8Decoder"}
METHODS = {"generate", "load_weights", "load_quantized_weights", "quantize_weights"}


def kind(node):
    """Literal values are kept (they are the model dimensions / paths / counts the callers use); anything else
    is reduced to its syntactic kind."""
    if isinstance(node, ast.Constant):
        return {"const": node.value}
    if isinstance(node, ast.List) and not node.elts:
        return {"kind": "empty_list"}
    return {"kind": type(node).__name__}


def scan(path, rel):
    with open(path) as f:
        tree = ast.parse(f.read(), rel)
    # names bound to `[]` somewhere in the file: the out-parameter lists
    empty_lists = {t.id for n in ast.walk(tree) if isinstance(n, ast.Assign) and isinstance(n.value, ast.List)
                   and not n.value.elts for t in n.targets if isinstance(t, ast.Name)}
    imports = sorted({f"{n.module}:{a.name}" for n in ast.walk(tree) if isinstance(n, ast.ImportFrom) and n.module
                      and n.module.startswith("decoder") for a in n.names})
    sites = []
    for n in ast.walk(tree):
        if not isinstance(n, ast.Call):
            continue
        f = n.func
        name = f.id if isinstance(f, ast.Name) else (f.attr if isinstance(f, ast.Attribute) else None)
        if name in CTORS or (name in METHODS and isinstance(f, ast.Attribute)):
            args = []
            for a in n.args:
                k = kind(a)
                if isinstance(a, ast.Name) and a.id in empty_lists:
                    k = {"kind": "out_list"}
                args.append(k)
            sites.append({"file": rel, "line": n.lineno, "callee": name, "args": args,
                          "kwargs": {k.arg: kind(k.value) for k in n.keywords}})
    return imports, sites


def main():
    out = {"imports": {}, "sites": []}
    for sub in ("api", "cli", "web"):
        for fn in sorted(os.listdir(os.path.join(REF, sub))):
            if fn.endswith(".py"):
                rel = f"{sub}/{fn}"
                imps, sites = scan(os.path.join(REF, rel), rel)
                if imps:
                    out["imports"][rel] = imps
                out["sites"] += sites
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {len(out['sites'])} call sites from {len(out['imports'])} files")
    return out


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
