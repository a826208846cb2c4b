#!/usr/bin/env python3
"""Embeds the headers the NVRTC translation unit of a user model includes into build/embedded_sources.inc
(C string literals), so that libcddp_b200.so carries its own kernel text."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = [("engine.h", "engine.h"), ("models.cuh", "models.cuh"), ("records.cuh", "records.cuh"), ("static_for.cuh", "static_for.cuh"),
         ("kernel_common.cuh", "kernel_common.cuh"), ("kernels_linearize.cuh", "kernels_linearize.cuh"),
         ("kernels_forward.cuh", "kernels_forward.cuh"), ("kernels_ipddp.cuh", "kernels_ipddp.cuh"), ("user_model.cuh", "user_model.cuh"),
         ("../../include/cddp_b200.h", "../../include/cddp_b200.h")]


def lit(text):
    out = []
    for line in text.splitlines():
        esc = line.replace("\\", "\\\\").replace('"', '\\"')
        out.append('"' + esc + '\\n"')
    return "\n".join(out) if out else '""'


def main():
    dst = sys.argv[1]
    parts = ["struct EmbeddedSource { const char *name; const char *text; };", "static const EmbeddedSource kEmbedded[] = {"]
    for path, name in FILES:
        with open(os.path.join(HERE, path)) as f:
            parts.append('{"%s",\n%s},' % (name, lit(f.read())))
    parts.append("};")
    parts.append("static const int kNumEmbedded = %d;" % len(FILES))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        f.write("\n".join(parts) + "\n")


if __name__ == "__main__":
    main()
