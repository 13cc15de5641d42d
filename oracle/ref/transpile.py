"""oracle/ref/transpile.py — TEST INFRASTRUCTURE ONLY.

Turns an assembled reference fragment shader (GLSL ES 3.00 text read from /root/reference by oracle/ref/assemble.py) into a
C++20 translation unit that compiles against oracle/ref/glsl_rt.h, and builds it into oracle/_ref/<name>_<hash>.so.

The transformation is purely syntactic — every statement, expression and constant of the shader is kept; only spellings that
differ between GLSL and C++ are changed:
  * float literals get an `f` suffix (GLSL literals are fp32);
  * `uniform` / `varying` / precision qualifiers are dropped (the globals become members of one struct per shader, copied per
    fragment so that GLSL's per-invocation globals keep their semantics), `layout(location = i) out` declares an output;
  * `inout T x` / `out T x` parameters become references, `in` disappears;
  * array constructors `T[n](...)` become make_arr<T>(...), `T[n] name;` becomes `T name[n];`
  * locals and globals declared without initialiser are value-initialised (the shaders read them; software GL
    zero-fills registers);
  * `discard` marks the fragment dead and `return` in main() freezes its outputs, but the invocation keeps running as a helper
    lane so that its quad neighbours can take derivatives (glsl_rt.h);
  * `a && fwidth(..)` evaluates both operands (as a SIMD rasteriser does), so every lane of a quad reaches the derivative;
  * `T x = x(...)` (GLSL: the new name is not yet in scope inside its initialiser) calls the member function explicitly.
No reference source is copied into the repository: the generated .cpp and .so live in oracle/_ref/ (git-ignored).
"""
from __future__ import annotations

import hashlib
import itertools
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
CXXFLAGS = ["-O2", "-std=c++20", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-fwrapv", "-w"]

SCALARS = {"float": ("f", 1), "int": ("i", 1), "uint": ("i", 1), "bool": ("b", 1)}
for _n in (2, 3, 4):
    SCALARS[f"vec{_n}"] = ("f", _n)
    SCALARS[f"ivec{_n}"] = ("i", _n)
    SCALARS[f"uvec{_n}"] = ("i", _n)
    SCALARS[f"mat{_n}"] = ("f", _n * _n)


def gen_swizzles(out_dir: str) -> None:
    for n in (2, 3, 4):
        lines = []
        for names in ("xyzw"[:n], "rgba"[:n]):
            for m in (2, 3, 4):
                for combo in itertools.product(range(n), repeat=m):
                    lines.append(f"swz<T, {n}, {', '.join(map(str, combo))}> {''.join(names[i] for i in combo)};")
        with open(os.path.join(out_dir, f"swz{n}.inc"), "w") as f:
            f.write("\n".join(lines) + "\n")


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


_FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def _match_brace(s: str, open_idx: int) -> int:
    depth = 0
    for i in range(open_idx, len(s)):
        if s[i] == "{":
            depth += 1
        elif s[i] == "}":
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced braces")


def preprocess(src: str) -> str:
    r = subprocess.run(["gcc", "-E", "-P", "-undef", "-nostdinc", "-x", "c", "-"], input=src, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("preprocessing the shader failed:\n" + r.stderr[:4000])
    return r.stdout


def drop_prototypes(s: str) -> str:
    """function prototypes at global scope (`T f(args);`): members of a class are visible before their definition"""
    out, depth, i, stmt_start = [], 0, 0, 0
    for i, ch in enumerate(s):
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                stmt_start = i + 1
        elif ch == ";" and depth == 0:
            stmt = s[stmt_start:i + 1]
            if re.fullmatch(r"\s*(?:const\s+)?\w+\s+\w+\s*\([^;{}=]*\)\s*;", stmt, flags=re.S):
                out.append((stmt_start, i + 1))
            stmt_start = i + 1
    for a, b in reversed(out):
        s = s[:a] + s[b:]
    return s


def parse_structs(src: str) -> dict:
    out = {}
    for m in re.finditer(r"\bstruct\s+(\w+)\s*\{([^}]*)\}", src):
        fields = []
        for decl in m.group(2).split(";"):
            toks = decl.replace(",", " ").split()
            toks = [t for t in toks if t not in ("highp", "mediump", "lowp")]
            if len(toks) >= 2:
                fields += [(toks[0], name) for name in toks[1:]]
        out[m.group(1)] = fields
    return out


def transpile(glsl: str, *, effect_main: bool = False) -> tuple[str, dict]:
    """returns (C++ struct body, info) — info: uniforms [(type, name)], outputs [name]"""
    s = glsl.replace("﻿", "")
    s = strip_comments(s)
    s = re.sub(r"^\s*precision\s+\w+\s+\w+\s*;", "", s, flags=re.M)
    s = re.sub(r"^\s*#version[^\n]*", "", s, flags=re.M)
    s = re.sub(r"^\s*#pragma[^\n]*", "", s, flags=re.M)
    s = re.sub(r"\b(highp|mediump|lowp)\b", "", s)
    s = preprocess(s)  # the GLSL preprocessor is the C preprocessor: #define / #if / function-like macros resolved here
    s = _FLOAT_LIT.sub(lambda m: m.group(1) + "f", s)
    s = drop_prototypes(s)

    # a swizzle of a swizzle (`c.rgb.rgb`, from macros like luminance(c) applied to `x.rgb`) is folded into one swizzle
    comp = {c: i for i, c in enumerate("xyzw")} | {c: i for i, c in enumerate("rgba")}

    field_names = {n for fields in parse_structs(s).values() for _, n in fields}  # InputTexel has a member called `rgb`

    def _fold(m):
        first, second = m.group(1), m.group(2)
        if first in field_names:
            return m.group(0)
        return "." + "".join(first[comp[c]] for c in second)

    prev = None
    while prev != s:
        prev = s
        s = re.sub(r"\.([xyzw]{2,4}|[rgba]{2,4})\.([xyzw]{1,4}|[rgba]{1,4})\b", _fold, s)

    structs = parse_structs(s)
    uniforms, outputs = [], []

    def _out(m):
        outputs.append((int(m.group(1)), m.group(3)))
        return f"{m.group(2)} {m.group(3)};"

    s = re.sub(r"layout\s*\(\s*location\s*=\s*(\d+)\s*\)\s*out\s+(\w+)\s+(\w+)\s*;", _out, s)

    def _uni(m):
        if (m.group(1), m.group(2)) in uniforms:
            return ""  # the same uniform declared twice (hbao.frag + the blue-noise chunk)
        uniforms.append((m.group(1), m.group(2)))
        return f"{m.group(1)} {m.group(2)};"

    s = re.sub(r"\buniform\s+(\w+)\s+(\w+)\s*;", _uni, s)
    s = re.sub(r"\bvarying\s+", "", s)
    # parameter qualifiers: arrays by reference; everything else through rt::io<T> (GLSL copy-in / copy-out, so that a swizzle
    # can be the argument), re-exposed under the parameter's own name as the first statement of the body
    s = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)\s*\[\s*(\w+)\s*\]", r"\1 (&\2)[\3]", s)

    def _fn(m):
        params, binds = m.group(3), []

        def _p(pm):
            binds.append(f"{pm.group(1)}& {pm.group(2)} = {pm.group(2)}_io_.v;")
            return f"io<{pm.group(1)}> {pm.group(2)}_io_"

        params = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", _p, params)
        return f"{m.group(1)} {m.group(2)}({params}) {{ " + " ".join(binds)

    s = re.sub(r"\b(\w+)\s+(\w+)\s*\(([^(){};]*\b(?:inout|out)\b[^(){};]*)\)\s*\{", _fn, s)
    s = re.sub(r"\bin\s+(?=\w)", "", s)
    # arrays
    s = re.sub(r"\b([A-Za-z_]\w*)\s*\[\s*\w*\s*\]\s*\(", r"make_arr<\1>(", s)
    s = re.sub(r"\b([A-Za-z_]\w*)\s+([A-Za-z_]\w*)\s*\[\s*(\w+)\s*\]\s*=\s*make_arr", r"std::array<\1, \3> \2 = make_arr", s)
    s = re.sub(r"\b([A-Za-z_]\w*)\s*\[\s*(\w+)\s*\]\s+([A-Za-z_]\w*)\s*;", r"\1 \3[\2]{};", s)

    # declarations without initialiser -> value-initialised
    def _zero(m):
        decls = [d.strip() for d in m.group(3).split(",")]
        return f"{m.group(1)}{m.group(2)} " + ", ".join(d + "{}" for d in decls) + ";"

    s = re.sub(r"(^|[;{}]\s*)(float|int|uint|bool)\s+([A-Za-z_][\w\s,\[\]]*);", _zero, s, flags=re.M)
    # T x = x(...)  ->  T x = this->x(...)
    s = re.sub(r"\b(\w+)\s+(\w+)\s*=\s*\2\s*\(", r"\1 \2 = this->\2(", s)
    s = re.sub(r"\bdiscard\s*;", "{ rt_discarded = true; }", s)
    # `a && fwidth(b) == 0.`: SIMD rasterisers evaluate both operands for every lane (no side effects), which is what makes the
    # derivative defined for the lanes whose `a` is false; C++'s short circuit would skip the site for them
    s = re.sub(r"&&(\s*(?:fwidth|dFdx|dFdy)\s*\()", r"&\1", s)

    # main(): `return;` freezes the outputs, the lane keeps running as a helper
    m = re.search(r"\bvoid\s+main\s*\(\s*\)\s*\{", s)
    if m:
        end = _match_brace(s, m.end() - 1)
        body = re.sub(r"\breturn\s*;", "{ rt_freeze(); }", s[m.end():end])
        s = s[:m.end()] + body + s[end:]
    elif not effect_main:
        raise ValueError("shader has no main()")

    outputs.sort()
    out_names = [n for _, n in outputs] or ["gl_FragColor"]
    if not outputs:
        s = "vec4 gl_FragColor;\n" + s
    info = dict(uniforms=uniforms, outputs=out_names, structs=structs)
    return s, info


def _setter_code(uniforms, structs) -> str:
    u_lines, t_lines = [], []

    def one(typ, path):
        if typ == "sampler2D":
            t_lines.append(f'    if (!std::strcmp(rtx_n, "{path}")) {{ {path}.t = rtx_t; return 0; }}')
        elif typ in SCALARS:
            kind, n = SCALARS[typ]
            u_lines.append(f'    if (!std::strcmp(rtx_n, "{path}")) return rtx_assign_{kind}(&{path}, {n}, rtx_p, rtx_count, {1 if typ == "bool" else 0});')
        elif typ in structs:
            for ft, fn in structs[typ]:
                one(ft, f"{path}.{fn}")
        else:
            raise ValueError(f"uniform of unsupported type {typ} {path}")

    for typ, name in uniforms:
        one(typ, name)
    return ("  int rtx_set_uniform(const char* rtx_n, const double* rtx_p, int rtx_count) {\n" + "\n".join(u_lines) + "\n    return -1;\n  }\n"
            "  int rtx_set_sampler(const char* rtx_n, const rt::TexDesc* rtx_t) {\n" + "\n".join(t_lines) + "\n    return -1;\n  }\n")


GLUE_HEAD = r"""
// GENERATED by oracle/ref/transpile.py from the reference's shader text — do not edit, do not commit.
#include "glsl_rt.h"
#include <cstdio>
#undef M_PI
namespace rtx_glue {
// uniform values arrive as doubles (what the reference's JS holds) and are converted the way gl.uniform* converts them
inline int rtx_assign_f(void* dst, int n, const double* p, int count, int) { if (count != n) return -2; float* d = (float*)dst; for (int k = 0; k < n; k++) d[k] = (float)p[k]; return 0; }
inline int rtx_assign_i(void* dst, int n, const double* p, int count, int) { if (count != n) return -2; int* d = (int*)dst; for (int k = 0; k < n; k++) d[k] = (int)(long long)p[k]; return 0; }
inline int rtx_assign_b(void* dst, int n, const double* p, int count, int) { if (count != n) return -2; *(bool*)dst = p[0] != 0.0; return 0; }
}
namespace rt {
using namespace rtx_glue;
struct Shader : FragBase {
  vec4 rtx_saved[8];
  void rt_freeze() { if (!rt_returned) { rtx_collect(rtx_saved); rt_returned = true; } }
  void rt_outputs(vec4* o) { if (rt_returned) { for (int k = 0; k < 8; k++) o[k] = rtx_saved[k]; } else rtx_collect(o); }
"""

GLUE_TAIL = r"""
};
}  // namespace rt
#undef main
extern "C" {
void* rtx_create() { return new rt::Shader(); }
void rtx_destroy(void* s) { delete (rt::Shader*)s; }
int rtx_uniform(void* s, const char* name, const double* v, int count) { return ((rt::Shader*)s)->rtx_set_uniform(name, v, count); }
void* rtx_tex_create(const void* data, int w, int h, int fmt, int linear, int repeat) {
  auto* t = new rt::TexDesc();
  t->base.data = data; t->base.w = w; t->base.h = h; t->base.fmt = fmt; t->base.linear = linear != 0; t->base.repeat = repeat != 0;
  return t;
}
void rtx_tex_add_mip(void* tex, const void* data, int w, int h) {
  auto* t = (rt::TexDesc*)tex;
  gl::Tex& l = t->mips.level[t->mips.levels++];
  l = t->base; l.data = data; l.w = w; l.h = h;
}
void rtx_tex_destroy(void* t) { delete (rt::TexDesc*)t; }
int rtx_sampler(void* s, const char* name, const void* tex) { return ((rt::Shader*)s)->rtx_set_sampler(name, (const rt::TexDesc*)tex); }
int rtx_run(void* s, int W, int H, int n_out, void* const* out_data, const int* out_fmt) {
  rt::run_fullscreen(*(const rt::Shader*)s, W, H, n_out, out_data, out_fmt);
  return 0;
}
int rtx_num_outputs() { return RTX_NUM_OUTPUTS; }
}
"""


def make_cpp(glsl: str, *, effect_main: bool = False) -> tuple[str, dict]:
    body, info = transpile(glsl, effect_main=effect_main)
    collect = "  void rtx_collect(vec4* o) { " + " ".join(f"o[{i}] = {n};" for i, n in enumerate(info["outputs"])) + " }\n"
    # rt_main is emitted inside the struct AFTER the shader text, so the shader's #defines apply to nothing of ours but `main`
    cpp = (GLUE_HEAD + body + "\n" + collect + "  void rt_main() { main(); }\n" + _setter_code(info["uniforms"], info["structs"]) +
           GLUE_TAIL.replace("RTX_NUM_OUTPUTS", str(len(info["outputs"]))))
    return cpp, info


def build(name: str, glsl: str, *, effect_main: bool = False, force: bool = False) -> tuple[str, dict]:
    """-> (path of the .so, info).  Cached by the hash of the generated C++ and of glsl_rt.h / glsl.h."""
    os.makedirs(OUT_DIR, exist_ok=True)
    cpp, info = make_cpp(glsl, effect_main=effect_main)
    h = hashlib.sha1()
    h.update(cpp.encode())
    for dep in (os.path.join(HERE, "glsl_rt.h"), os.path.join(os.path.dirname(HERE), "glsl.h"), __file__):
        with open(dep, "rb") as f:
            h.update(f.read())
    tag = h.hexdigest()[:12]
    so = os.path.join(OUT_DIR, f"{name}_{tag}.so")
    if force or not os.path.exists(so):
        if not os.path.exists(os.path.join(OUT_DIR, "swz4.inc")):
            gen_swizzles(OUT_DIR)
        src = os.path.join(OUT_DIR, f"{name}_{tag}.cpp")
        with open(src, "w") as f:
            f.write(cpp)
        cmd = [os.environ.get("RFX_HOST_CXX", "g++"), *CXXFLAGS, "-I", HERE, "-I", OUT_DIR, "-shared", "-o", so + ".tmp", src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for the transpiled shader {name} ({src}):\n{r.stderr[:6000]}")
        os.replace(so + ".tmp", so)
        if not os.environ.get("RFX_KEEP_GENERATED"):
            os.remove(src)  # the generated C++ carries the reference's shader text: not kept around
    return so, info
