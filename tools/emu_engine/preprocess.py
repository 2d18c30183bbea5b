"""
csrc/symeig.cu -> plain C++ for the host (see emu_cuda.h).  Only textual rewrites of the constructs a host compiler
cannot take; the algorithmic source is untouched:
  * `kernel<<<grid, block, smem, stream>>>(args);`      -> emu_launch(grid, block, smem, emu_bind(kernel, args));
  * `cudaLaunchCooperativeKernel(fn, g, b, kargs, ...)` -> emu_coop<PostArgs>(fn, g, b, kargs, smem)
  * `__shared__ T x[N];` / `extern __shared__ T x[];`   -> per-CTA arena / dynamic buffer of the emulated CTA
  * the four functions whose bodies are PTX (`gtimer`, `cp_async16`, `cp_async_wait_all`, `fast_rcp`) are dropped: their
    stand-ins live in emu_cuda.h; the PTX fences of the grid barrier become std::atomic fences.

    python tools/emu_engine/preprocess.py <symeig.cu> <out.cpp>
"""
import re
import sys


def _drop_function(src: str, head: str) -> str:
    i0 = src.index(head)
    line_end = src.index("\n", i0)
    if src[i0:line_end].rstrip().endswith("}"):                 # one-liner
        return src[:i0] + src[line_end + 1:]
    i1 = src.index("\n}\n", i0) + 3
    return src[:i0] + src[i1:]


def _balanced(src: str, start: int, open_ch: str, close_ch: str) -> int:
    """index just after the bracket that closes the one at `start`"""
    depth = 0
    for i in range(start, len(src)):
        if src[i] == open_ch:
            depth += 1
        elif src[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def _rewrite_launches(src: str) -> str:
    out, pos = [], 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            break
        # kernel name (with optional template arguments and namespace) right before <<<
        j = i
        if src[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] in "_:"):
            j -= 1
        name = src[j:i]
        cfg_end = src.index(">>>", i)
        cfg = [c.strip() for c in _split_args(src[i + 3:cfg_end])]
        grid, block, smem = cfg[0], cfg[1], (cfg[2] if len(cfg) > 2 else "0")
        a0 = cfg_end + 3
        assert src[a0] == "(", src[a0 - 20:a0 + 20]
        a1 = _balanced(src, a0, "(", ")")
        args = src[a0:a1]
        assert src[a1] == ";"
        out.append(src[pos:j])
        # arguments are evaluated HERE, in the launching thread (as a real launch does), not inside the worker threads
        out.append("emu_launch(dim3(%s), dim3(%s), (size_t)(%s), emu_bind(%s, %s));"
                   % (grid, block, smem, name, args[1:-1]))
        pos = a1 + 1
    return "".join(out)


def _split_args(s: str):
    parts, depth, cur = [], 0, ""
    for i, ch in enumerate(s):
        arrow = ch == ">" and i > 0 and s[i - 1] == "-"          # `->` is not a closing bracket
        if ch in "(<[":
            depth += 1
        elif ch in ")>]" and not arrow:
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def transform(src: str, include_dir: str = None) -> str:
    src = src.replace('#include "matvec.cuh"', '#include "emu_cuda.h"').replace('#include "common.cuh"',
                                                                                 '#include "emu_cuda.h"')
    m = re.search(r'#include "(solve_common\.cuh)"', src)
    if m:                                                       # the solvers' shared header is inlined, rewritten too
        import os
        inc = open(os.path.join(include_dir, m.group(1))).read().replace("#pragma once", "")
        src = src.replace(m.group(0), transform(inc, include_dir))
    for head in ("__device__ __forceinline__ unsigned long long gtimer() {",
                 "__device__ __forceinline__ void cp_async16(void* dst, const void* src) {",
                 "__device__ __forceinline__ void cp_async_wait_all() {",
                 "__device__ __forceinline__ double fast_rcp(double x) {"):
        if head in src:
            src = _drop_function(src, head)
    src = src.replace('asm volatile("fence.acq_rel.gpu;" ::: "memory");', "emu_fence();")
    src = re.sub(r"extern __shared__ (?:__align__\(\d+\) )?(\w+(?: \w+)?) (\w+)\[\];",
                 r"\1* \2 = emu_dyn_smem<\1>();", src)
    src = re.sub(r"__shared__ (\w+) (\w+)\[([^\]]+)\]\[([^\]]+)\];",
                 r"\1 (*\2)[\4] = reinterpret_cast<\1 (*)[\4]>(emu_shared<\1>(__COUNTER__, (\3) * (\4)));", src)
    src = re.sub(r"__shared__ (\w+) (\w+)\[([^\]]+)\];", r"\1* \2 = emu_shared<\1>(__COUNTER__, \3);", src)
    src = re.sub(r"__shared__ (\w+) (\w+);", r"\1& \2 = *emu_shared<\1>(__COUNTER__, 1);", src)
    src = re.sub(r"cudaLaunchCooperativeKernel\(po_fn, dim3\(po_grid\), dim3\(PO_THREADS\), kargs, (\w+), st\)",
                 r"emu_coop<PostArgs>(po_fn, dim3(po_grid), dim3(PO_THREADS), kargs, \1)", src)
    src = re.sub(r"cudaLaunchCooperativeKernel\(sh_fn, dim3\(sh_grid\), dim3\(PO_THREADS\), kargs, (\w+), st\)",
                 r"emu_coop<ShardArgs>(sh_fn, dim3(sh_grid), dim3(PO_THREADS), kargs, \1)", src)
    src = _rewrite_launches(src)
    assert "asm" not in re.sub(r"//.*", "", src), "PTX left in the host source"
    assert "__shared__" not in src and "<<<" not in src
    return src


if __name__ == "__main__":
    import os
    open(sys.argv[2], "w").write(transform(open(sys.argv[1]).read(), os.path.dirname(os.path.abspath(sys.argv[1]))))
