"""Extract the reference CLIs (flag, type, default, action, choices) with ast: /root/reference/distill.py:625-679 ->
tests/golden/cli.json, /root/reference/buffer.py:119-160 -> tests/golden/cli_buffer.json.

    python tests/golden/make_cli_golden.py      # build container only
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/distill.py"


def lit(node):
    try:
        return ast.literal_eval(node)
    except Exception:
        return "<non-literal>"


def main(SRC=SRC, out="cli.json", source="distill.py:625-679"):
    tree = ast.parse(open(SRC).read())
    flags = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            name = lit(node.args[0])
            if not (isinstance(name, str) and name.startswith("--")):
                continue
            entry = {"flag": name, "line": node.lineno}
            for kw in node.keywords:
                if kw.arg == "type":
                    entry["type"] = kw.value.id if isinstance(kw.value, ast.Name) else "<expr>"
                elif kw.arg in ("default", "action", "choices"):
                    entry[kw.arg] = lit(kw.value)
            flags.append(entry)
    flags.sort(key=lambda e: e["line"])
    with open(os.path.join(HERE, out), "w") as f:
        json.dump({"source": source, "flags": flags}, f, indent=1)
    print(out, len(flags), "flags")


if __name__ == "__main__":
    main()
    main("/root/reference/buffer.py", "cli_buffer.json", "buffer.py:119-160")
