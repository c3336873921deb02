"""TEST INFRASTRUCTURE ONLY.  Rewrites the CUDA launch syntax of a product .cu file so that g++ can compile the file --
host code included -- against tests/emu/cuda_emu.h:

    kernel<targs><<<grid, block, smem, stream>>>(args);   ->   ::cuda_emu::launch_dim(grid, block, [&] { kernel<targs>(args); });

Nothing else is touched: the host logic that sizes grids, allocates scratch and orders the launches is the product's.
"""
import re
import sys


def _match_back(src: str, i: int, open_ch: str, close_ch: str) -> int:
    """src[i] == close_ch: index of the matching open_ch."""
    depth = 0
    while i >= 0:
        c = src[i]
        if c == close_ch:
            depth += 1
        elif c == open_ch:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced %s%s" % (open_ch, close_ch))


def _match_fwd(src: str, i: int, open_ch: str, close_ch: str) -> int:
    """src[i] == open_ch: index of the matching close_ch."""
    depth = 0
    while i < len(src):
        c = src[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced %s%s" % (open_ch, close_ch))


def _split_top(s: str):
    out, depth, cur = [], 0, []
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    out.append("".join(cur).strip())
    return out


def rewrite_launches(src: str) -> str:
    out, pos, count = [], 0, 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            break
        # kernel expression: identifier, optionally followed by <template arguments>
        k_end = i
        j = i - 1
        while src[j].isspace():
            j -= 1
        if src[j] == ">":
            j = _match_back(src, j, "<", ">") - 1
        while re.match(r"[A-Za-z0-9_:]", src[j]):
            j -= 1
        k_start = j + 1
        kernel = src[k_start:k_end].strip()
        close = src.find(">>>", i)
        cfg = _split_top(src[i + 3:close])
        assert len(cfg) in (2, 3, 4), cfg
        a = close + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", src[a:a + 40]
        a_end = _match_fwd(src, a, "(", ")")
        args = src[a + 1:a_end]
        out.append(src[pos:k_start])
        out.append("::cuda_emu::launch_dim(%s, %s, [&] { %s(%s); })" % (cfg[0], cfg[1], kernel, args))
        pos = a_end + 1
        count += 1
    out.append(src[pos:])
    return "".join(out), count


if __name__ == "__main__":
    text, n = rewrite_launches(open(sys.argv[1]).read())
    open(sys.argv[2], "w").write(text)
    print("%d launches rewritten" % n)
