"""TEST INFRASTRUCTURE ONLY: the two mechanical source rewrites that let g++ compile the engine's .cu files over
tests/harness/cuda_on_cpu/cuda_runtime.h.

1. `kernel<<<grid, block, smem, stream>>>(args);`  ->  `::cuda_on_cpu::launch((grid), (block), [&]() { kernel(args); });`
2. every inline-PTX statement `asm [volatile]("..." : outputs : inputs : clobbers);` -> a call of the C++ function in
   cuda_runtime.h (namespace cuda_on_cpu::ptx) that states what that PTX instruction does. Only the handful of
   instruction forms the traversal kernels use are known; anything else becomes `::cuda_on_cpu::ptx::unsupported("...")`,
   which aborts if it is ever executed (so a file compiles even when a path of it -- e.g. the TMA staging of the
   reference-format kernels -- cannot run here).

Everything else in the file -- control flow, arithmetic, warp votes, the stack discipline -- is compiled as written."""
from __future__ import annotations

import re

DYNAMIC_SHARED = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\[\];")
LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)<<<(.*?)>>>\((.*?)\);", re.S)


def split_top_level(text: str, sep: str = ",") -> list[str]:
    parts, depth, cur, in_str, prev = [], 0, "", False, ""
    for ch in text:
        if in_str:
            cur += ch
            if ch == '"' and prev != "\\":
                in_str = False
        elif ch == '"':
            in_str = True
            cur += ch
        elif ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            depth += ch in "([{"
            depth -= ch in ")]}"
            cur += ch
        prev = ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(source: str) -> tuple[str, int]:
    def repl(m):
        cfg = split_top_level(m.group(2))
        assert len(cfg) in (2, 3, 4), m.group(0)
        dynamic = f", ({cfg[2]})" if len(cfg) >= 3 and cfg[2] != "0" else ""
        return f"::cuda_on_cpu::launch(({cfg[0]}), ({cfg[1]}), [&]() {{ {m.group(1)}({m.group(3)}); }}{dynamic});"
    text, n = LAUNCH.subn(repl, source)
    # dynamic shared memory: `extern __shared__ [__align__(n)] T name[];` -> a pointer to the launch's dynamic block
    text = DYNAMIC_SHARED.sub(lambda m: f"{m.group(1)}* {m.group(2)} = static_cast<{m.group(1)}*>(::cuda_on_cpu::dynamic_shared());", text)
    return text, n


# normalised PTX template -> C++ (operands substituted for %N)
_MEM = r"\[%(\d+)(?:\+(\d+))?\]"
PTX_FORMS: list[tuple[re.Pattern, str]] = [
    (re.compile(r"^mov\.b64 %(\d+), \{%(\d+),%(\d+)\};$"), "::cuda_on_cpu::ptx::pack2({0}, {1}, {2});"),
    (re.compile(r"^mov\.b64 \{%(\d+),%(\d+)\}, %(\d+);$"), "::cuda_on_cpu::ptx::unpack2({0}, {1}, {2});"),
    (re.compile(r"^mov\.b64 %(\d+), %(\d+);$"), "{0} = {1};"),
    (re.compile(r"^fma\.rn\.ftz\.f32x2 %(\d+), %(\d+), %(\d+), %(\d+);$"), "::cuda_on_cpu::ptx::fma2({0}, {1}, {2}, {3});"),
    (re.compile(r"^max\.ftz\.f32 %(\d+), %(\d+), %(\d+), %(\d+);$"), "{0} = fmaxf(fmaxf({1}, {2}), {3});"),
    (re.compile(r"^min\.ftz\.f32 %(\d+), %(\d+), %(\d+), %(\d+);$"), "{0} = fminf(fminf({1}, {2}), {3});"),
    (re.compile(r"^mad\.wide\.u32 %(\d+), %(\d+), (\d+), %(\d+);$"), "{0} = (unsigned long long)(uint32_t)({1}) * {imm}ull + (unsigned long long)({3});"),
    (re.compile(r"^ld\.global\.nc\.v8\.f32 \{%0,%1,%2,%3,%4,%5,%6,%7\}, " + _MEM + r";$"), "LD8F"),
    (re.compile(r"^ld\.global\.nc\.v8\.u32 \{%0,%1,%2,%3,%4,%5,%6,%7\}, " + _MEM + r";$"), "LD8U"),
    (re.compile(r"^ld\.v4\.f32 \{%0,%1,%2,%3\}, " + _MEM + r";$"), "LD4F"),
    (re.compile(r"^ld\.global\.nc\.v4\.b64 \{%0,%1,%2,%3\}, " + _MEM + r";$"), "LD4Q"),
    (re.compile(r"^st\.local\.u32 \[%(\d+)\], %(\d+);$"), "*::cuda_on_cpu::ptx::local_word({0}) = {1};"),
    (re.compile(r"^ld\.local\.u32 %(\d+), \[%(\d+)\];$"), "{0} = *::cuda_on_cpu::ptx::local_word({1});"),
    (re.compile(r"^st\.shared\.u32 \[%(\d+)\], %(\d+);$"), "*::cuda_on_cpu::ptx::shared_word({0}) = {1};"),
    (re.compile(r"^ld\.shared\.u32 %(\d+), \[%(\d+)\];$"), "{0} = *::cuda_on_cpu::ptx::shared_word({1});"),
    (re.compile(r"^cvta\.shared\.u64 %(\d+), %(\d+);$"), "{0} = (unsigned long long)(uintptr_t)::cuda_on_cpu::ptx::shared_pointer({1});"),
    # PlainStack::pushIf: a predicated store + bump
    (re.compile(r"^\{ \.reg \.pred pu; setp\.ne\.u32 pu, %2, 0; @pu st\.local\.u32 \[%1\], %3; @pu add\.u32 %0, %0, 4; \}$"),
     "if ({2}) {{ *::cuda_on_cpu::ptx::local_word({1}) = {3}; {0} += 4; }}"),
]


def _find_asm_statements(source: str):
    """Yields (start, end, inner) for every `asm [volatile]( inner );`."""
    for m in re.finditer(r"\basm\s*(?:volatile\s*)?\(", source):
        i, depth, in_str, prev = m.end(), 1, False, ""
        while depth:
            ch = source[i]
            if in_str:
                if ch == '"' and prev != "\\":
                    in_str = False
            elif ch == '"':
                in_str = True
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            prev = ch
            i += 1
        j = i
        while source[j] in " \t\n":
            j += 1
        assert source[j] == ";", source[m.start():j + 1]
        yield m.start(), j + 1, source[m.end():i - 1]


def _operand_exprs(section: str) -> list[str]:
    out = []
    for op in split_top_level(section):
        if not op:
            continue
        m = re.match(r'^"[^"]*"\s*\((.*)\)$', op, re.S)
        assert m, op
        out.append("(" + m.group(1).strip() + ")")
    return out


def translate_asm(inner: str) -> tuple[str, bool]:
    sections = split_top_level(inner, ":")
    literals = re.findall(r'"((?:[^"\\]|\\.)*)"', sections[0])
    template = " ".join("".join(literals).replace("\\n", " ").replace("\\t", " ").split())
    ops = []
    for s in sections[1:3]:
        ops += _operand_exprs(s)
    for pattern, code in PTX_FORMS:
        m = pattern.match(template)
        if not m:
            continue
        if code in ("LD8F", "LD8U", "LD4Q", "LD4F"):
            n = 8 if code in ("LD8F", "LD8U") else 4
            addr, off = ops[int(m.group(1))], m.group(2) or "0"
            fn = code.lower()
            return f"::cuda_on_cpu::ptx::{fn}((unsigned long long){addr} + {off}ull, " + ", ".join(ops[:n]) + ");", True
        if "{imm}" in code:
            return code.format(ops[int(m.group(1))], ops[int(m.group(2))], None, ops[int(m.group(4))], imm=m.group(3)), True
        if pattern.pattern.startswith(r"^\{ \.reg"):
            return code.format(*ops), True
        return code.format(*[ops[int(g)] for g in m.groups()]), True
    escaped = template.replace("\\", "\\\\").replace('"', '\\"')
    return f'::cuda_on_cpu::ptx::unsupported("{escaped}");', False


def rewrite_asm(source: str) -> tuple[str, int, list[str]]:
    """Returns (source, translated statements, templates left unsupported)."""
    out, last, done, unsupported = [], 0, 0, []
    for start, end, inner in _find_asm_statements(source):
        code, ok = translate_asm(inner)
        out.append(source[last:start])
        out.append(code)
        last = end
        if ok:
            done += 1
        else:
            unsupported.append(code)
    out.append(source[last:])
    return "".join(out), done, unsupported
