"""TEST INFRASTRUCTURE: compile a plain-SIMT .cu file of genesis_b200/csrc for the CPU emulation of tests/cuda_emu/cuda_emu.h.

The source is used as it is except for the launch statements, which g++ cannot parse:
    kernel<T...><<<grid, block, smem, stream>>>(args);   ->   emu::launch(grid, block, [&] { kernel<T...>(args); });
so the extern "C" host entry points (argument checks, grid computation, kernel selection) are exercised too."""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), 'genesis_b200', 'csrc')
OUT = os.path.join(HERE, '_build')

PRELUDE = '''#define CUDA_EMU_MAIN
#include <cstring>
#include "cuda_emu.h"
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaMemset2DAsync(void* p, size_t pitch, int v, size_t w, size_t h, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) std::memset((char*)p + r * pitch, v, w);
    return cudaSuccess;
}
'''


def _balanced(text, i, open_ch, close_ch):
    """text[i] == open_ch -> index just past the matching close_ch."""
    depth = 0
    for j in range(i, len(text)):
        if text[j] == open_ch:
            depth += 1
        elif text[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError('unbalanced')


def _split_top(s):
    parts, depth, cur = [], 0, ''
    for ch in s:
        if ch in '(<[':
            depth += 1
        elif ch in ')>]':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur.strip())
            cur = ''
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


# ---- inline PTX -> calls into tests/cuda_emu/cuda_emu_sm100.h
PTX_RULES = [   # (substring of the PTX text, C++ statement; {o0} = first output operand, {i0}.. = input operands)
    ('elect.sync', '{o0} = emu::elect_one();'),
    # round to nearest (ties away from zero in magnitude) to a 10-bit mantissa, kept in an fp32 container
    ('cvt.rna.tf32.f32', '{{ uint32_t b__; float f__ = {i0}; std::memcpy(&b__, &f__, 4); b__ = (b__ + 0x1000u) & 0xFFFFE000u; {o0} = b__; }}'),
    ('mbarrier.init', 'emu::mbar_init({i0}, {i1});'),
    ('mbarrier.arrive.expect_tx', 'emu::mbar_expect_tx({i0}, {i1});'),
    ('mbarrier.arrive', 'emu::mbar_arrive({i0});'),
    ('mbarrier.try_wait.parity', '{o0} = emu::mbar_try_wait({i0}, {i1});'),
    ('cp.async.bulk.tensor.4d', 'emu::tma_load({i0}, (const CUtensorMap*)({i1}), {i2}, 4, {i3}, {i4}, {i5}, {i6});'),
    ('cp.async.bulk.tensor.3d', 'emu::tma_load({i0}, (const CUtensorMap*)({i1}), {i2}, 3, {i3}, {i4}, {i5}, 0);'),
    ('cp.async.bulk', 'emu::unsupported("cp.async.bulk (bulk-copy epilogue)");'),
    ('prefetch.tensormap', ';'),
    ('tcgen05.mma', 'emu::mma_tf32({i0}, {i1}, {i2}, {i3}, {i4});'),
    ('tcgen05.commit', 'emu::mbar_arrive({i0});'),
    ('tcgen05.alloc', 'emu::tmem_alloc({i0});'),
    ('tcgen05.relinquish', ';'), ('tcgen05.dealloc', ';'), ('tcgen05.fence', ';'), ('tcgen05.wait', ';'),
    ('tcgen05.ld', 'emu::tmem_ld32({i0}, &({o0}));'),
    ('fence.', ';'),
    ('bar.sync 1, 128', 'emu::named_barrier(1, 128);'),
    ('bar.sync 1, %0', 'emu::named_barrier(1, {i0});'),
    ('%%smid', '{o0} = 0;'),
]


def _asm_sections(body):
    """body of asm(...) -> (ptx text, [output exprs], [input exprs])"""
    i, ptx = 0, ''
    n = len(body)
    while True:                       # adjacent string literals
        while i < n and body[i].isspace():
            i += 1
        if i < n and body[i] == '"':
            j = i + 1
            while body[j] != '"' or body[j - 1] == '\\':
                j += 1
            ptx += body[i + 1:j]
            i = j + 1
        else:
            break
    rest = body[i:]
    sections, depth, cur, in_str = [], 0, '', False
    for ch in rest:
        if ch == '"':
            in_str = not in_str
        if not in_str:
            if ch in '([':
                depth += 1
            elif ch in ')]':
                depth -= 1
            if ch == ':' and depth == 0:
                sections.append(cur)
                cur = ''
                continue
        cur += ch
    sections.append(cur)
    sections = sections[1:]            # text before the first ':' is empty

    def operands(sec):
        out = []
        for m in re.finditer(r'"[^"]*"\s*\(', sec):
            a = m.end() - 1
            b = _balanced(sec, a, '(', ')')
            out.append(sec[a + 1:b - 1].strip())
        return out
    outs = operands(sections[0]) if len(sections) > 0 else []
    ins = operands(sections[1]) if len(sections) > 1 else []
    return ptx, outs, ins


def translate_asm(src):
    out, pos = '', 0
    for m in re.finditer(r'\basm\s*(?:volatile)?\s*\(', src):
        if m.start() < pos:
            continue
        a = m.end() - 1
        b = _balanced(src, a, '(', ')')
        assert src[b] == ';', src[m.start():b + 5]
        ptx, outs, ins = _asm_sections(src[a + 1:b - 1])
        for key, stmt in PTX_RULES:
            if key in ptx:
                fmt = {('o%d' % i): e for i, e in enumerate(outs)}
                fmt.update({('i%d' % i): e for i, e in enumerate(ins)})
                repl = stmt.format(**fmt)
                break
        else:
            raise ValueError('no emulation rule for PTX: %s' % ptx[:80])
        out += src[pos:m.start()] + '{ ' + repl + ' }'
        pos = b + 1
    return out + src[pos:]


def inline_includes(src, seen):
    """#include "umma.cuh" -> its text, once per translation unit (it holds inline PTX that must be translated too)"""
    def sub(m):
        name = m.group(1)
        if name == 'common.cuh':
            return m.group(0)
        if name in seen:
            return ''
        seen.add(name)
        text = open(os.path.join(CSRC, name)).read().replace('#pragma once', '')
        return inline_includes(text, seen)
    return re.sub(r'#include\s+"(\w+\.cuh)"', sub, src)


def transform(src, seen=None):
    src = translate_asm(inline_includes(src, set() if seen is None else seen))
    out, pos = '', 0
    for m in re.finditer(r'<<<', src):
        i = m.start()
        if i < pos:
            continue
        # kernel name (with template arguments) = the token run before <<<
        k = i
        depth = 0
        while k > 0:
            ch = src[k - 1]
            if ch == '>':
                depth += 1
            elif ch == '<':
                depth -= 1
            elif depth == 0 and not (ch.isalnum() or ch in '_:'):
                break
            k -= 1
        name = src[k:i]
        j = src.index('>>>', i)
        cfg = _split_top(src[i + 3:j])
        a0 = j + 3
        assert src[a0] == '(', src[a0:a0 + 20]
        a1 = _balanced(src, a0, '(', ')')
        semi = src[a1] == ';'           # a launch inside a macro body has no semicolon of its own
        out += src[pos:k] + 'emu::launch(%s, %s, [&] { %s%s; })%s' % (cfg[0], cfg[1], name, src[a0:a1], ';' if semi else '')
        pos = a1 + (1 if semi else 0)
    out += src[pos:]
    # dynamic shared memory: one static arena (blocks run one after another)
    return re.sub(r'extern\s+__shared__\s+(\w+)\s+(\w+)\[\];', r'\1* \2 = reinterpret_cast<\1*>(emu::dyn_smem);', out)


# files whose extern "C" entry points call into another file: built into one shared object
LINK_WITH = {'igemm_tc.cu': ['igemm_halo.cu']}


def build(cu_name):
    """-> path of the shared object emulating genesis_b200/csrc/<cu_name> (cached on the source hash)."""
    src = open(os.path.join(CSRC, cu_name)).read()
    seen = set()
    code = PRELUDE + transform(src, seen)
    for other in LINK_WITH.get(cu_name, []):            # same translation unit: one emulator state, shared headers once
        code += '\n// ---- linked: %s\n' % other + transform(open(os.path.join(CSRC, other)).read(), seen)
    deps = code + ''.join(open(os.path.join(HERE, h)).read() for h in ('cuda_emu.h', 'cuda_emu_sm100.h')) + \
        open(os.path.join(CSRC, 'common.cuh')).read()
    tag = hashlib.sha1(deps.encode()).hexdigest()[:12]
    os.makedirs(OUT, exist_ok=True)
    base = os.path.join(OUT, '%s_%s' % (cu_name[:-3], tag))
    so = base + '.so'
    if not os.path.exists(so):
        with open(base + '.cpp', 'w') as f:
            f.write(code)
        cmd = ['g++', '-O1', '-std=c++17', '-pthread', '-shared', '-fPIC', '-w', '-I', os.path.join(HERE, 'shim'), '-I', HERE,
               '-I', CSRC, base + '.cpp', '-o', so]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('g++ failed for the emulation of %s:\n%s' % (cu_name, r.stderr[-4000:]))
    return so
