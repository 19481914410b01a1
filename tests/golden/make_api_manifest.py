"""Generates tests/golden/reference_api_manifest.json: for every module under the reference's src/ package, the public
top-level names it defines (classes, functions) or re-exports (`__all__` of the package __init__ files).  Parsed with
`ast` - nothing of the reference is imported or copied, only names.  Run in the build container (the reference tree
does not exist on the GPU box):  python tests/golden/make_api_manifest.py /root/reference"""
import ast
import json
import os
import sys


def public_names(path):
    tree = ast.parse(open(path).read())
    names = []
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and not node.name.startswith("_"):
            names.append(node.name)
        elif isinstance(node, ast.Assign):
            for target in node.targets:
                if isinstance(target, ast.Name) and target.id == "__all__":
                    names.extend(ast.literal_eval(node.value))
    return sorted(set(names))


def main(root):
    src = os.path.join(root, "src")
    manifest = {}
    for directory, _, files in os.walk(src):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(directory, f), src)
            module = rel[:-3].replace(os.sep, ".")
            if module.endswith("__init__"):
                module = module[: -len("__init__")].rstrip(".")
            manifest[module] = public_names(os.path.join(directory, f))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_api_manifest.json")
    json.dump(manifest, open(out, "w"), indent=1, sort_keys=True)
    print(f"{len(manifest)} modules, {sum(len(v) for v in manifest.values())} names -> {out}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
