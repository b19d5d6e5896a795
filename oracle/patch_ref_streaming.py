#!/usr/bin/env python
"""Writes syntax-patched COPIES of two reference headers into oracle/_ref/patched/ (git-ignored
build output) so that the reference's own streamingSK kernel can be instantiated by nvcc + gcc.

TEST INFRASTRUCTURE ONLY.  Nothing of the reference is stored in this repository: the script
holds token-level edits only and reads the sources where they lie under the reference checkout.
The edits change no behaviour -- they are what MSVC accepts silently and gcc / EDG reject
(SURVEY.md section 8(c)):
  * `typename Load load;`                      -> `Load load;` (typename on a non-dependent name)
  * `X::TempStorage` inside class templates    -> `typename X::TempStorage`
  * explicit specialisations of the member template `compact<VARIANT>` at class scope
                                               -> overloads on std::integral_constant + a forwarding template
usage: patch_ref_streaming.py <reference implementation/src dir> <output dir>
"""
import os
import re
import sys


def patch_streaming(src: str) -> str:
    n = 0

    def sub(pattern, repl, text, count=0, flags=0):
        nonlocal n
        out, k = re.subn(pattern, repl, text, count=count, flags=flags)
        if k == 0:
            raise SystemExit(f"patch_ref_streaming: pattern not found: {pattern!r}")
        n += k
        return out

    s = src
    s = sub(r"typename\s+Load\s+load;", "Load load;", s)
    s = sub(r"(?<!typename )\b(BlockLoad(?:Float3|Float4|Uint)T)::TempStorage", r"typename \1::TempStorage", s)
    # primary member template declaration -> forwarding template over tag-dispatched overloads
    s = sub(r"template\s*<Variant\s+VARIANT>\s*__device__\s+__forceinline__\s+void\s+compact\(\s*AABB&\s*aabb,\s*"
            r"Thread\s*\(&thread\)\[ITEMS_PER_THREAD\],\s*Threads\s+d_threads,\s*int&\s*n_active\);",
            "template <Variant V_>\n  __device__ __forceinline__ void compact(AABB& aabb, Thread (&thread)[ITEMS_PER_THREAD],\n"
            "                                          Threads d_threads, int& n_active) {\n"
            "    compact_impl(std::integral_constant<Variant, V_>(), aabb, thread, d_threads, n_active);\n  }", s)
    s = sub(r"template\s*<>\s*__device__\s+__forceinline__\s+void\s+compact<(kClassic|kSortingRays)>\(",
            r"__device__ __forceinline__ void compact_impl(std::integral_constant<Variant, \1>, ", s)
    s = s.replace("#include <cub/cub.cuh>", "#include <cub/cub.cuh>\n#include <type_traits>", 1)
    return s


def patch_morton(src: str) -> str:
    out, k = re.subn(r"typedef\s+BlockRadixSortT::TempStorage\s+TempStorage;",
                     "typedef typename BlockRadixSortT::TempStorage TempStorage;", src)
    if k != 1:
        raise SystemExit("patch_ref_streaming: MortonSort.h pattern not found")
    return out


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for name, fn in (("StreamingVolPTsk_kernel.cuh", patch_streaming), ("MortonSort.h", patch_morton)):
        with open(os.path.join(ref, name)) as f:
            text = fn(f.read())
        with open(os.path.join(out, name), "w") as f:
            f.write(text)


if __name__ == "__main__":
    main()
