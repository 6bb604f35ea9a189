"""A small SPIR-V interpreter: enough of the instruction set to EXECUTE the reference's precompiled compute shader
`assets/probe_conv_comp.spv` (the "IrradianceProbeConvolve" pass) one invocation at a time on the CPU.

It is test infrastructure for generating golden vectors (tests/golden/make_probe_conv_golden.py): the instructions that run
are the reference's own compiled code; what this file defines is only what the SPIR-V / Vulkan specs leave to the
implementation — fp32 arithmetic is IEEE single (numpy.float32), OpDot sums left to right, GLSL.std.450 Normalize is
v * (1 / sqrt(dot(v, v))), integer division truncates — the same choices DESIGN.md states for the oracle.

Supported: the opcodes that module uses (see OPS below); anything else raises NotImplementedError.
"""
import struct

import numpy as np

f32 = np.float32


class Ptr:
    """Pointer = (root list, path of indices); variables live in one-element root lists."""

    def __init__(self, root, path=()):
        self.root, self.path = root, tuple(path)

    def load(self):
        v = self.root
        for i in self.path:
            v = v[i]
        return v

    def store(self, val):
        v = self.root
        for i in self.path[:-1]:
            v = v[i]
        v[self.path[-1]] = val

    def chain(self, idx):
        return Ptr(self.root, self.path + tuple(idx))


def _u32(x):
    return int(x) & 0xFFFFFFFF


def _s32(x):
    x = int(x) & 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def _map(fn, *a):
    if isinstance(a[0], list):
        return [fn(*[x[i] if isinstance(x, list) else x for x in a]) for i in range(len(a[0]))]
    return fn(*a)


class Module:
    def __init__(self, path):
        data = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(data) // 4), data)
        assert w[0] == 0x07230203
        self.types, self.consts, self.names, self.member_names, self.decor = {}, {}, {}, {}, {}
        self.globals, self.functions, self.entry, self.global_init = {}, {}, None, {}
        i, cur = 5, None
        while i < len(w):
            op, wc = w[i] & 0xFFFF, w[i] >> 16
            a = w[i + 1:i + wc]
            if op == 5:
                self.names[a[0]] = self._str(a[1:])
            elif op == 6:
                self.member_names[(a[0], a[1])] = self._str(a[2:])
            elif op == 15:
                self.entry = a[1]
            elif op == 71:
                self.decor.setdefault(a[0], []).append(a[1:])
            elif op == 19:
                self.types[a[0]] = ("void",)
            elif op == 20:
                self.types[a[0]] = ("bool",)
            elif op == 21:
                self.types[a[0]] = ("int", a[1], a[2])
            elif op == 22:
                self.types[a[0]] = ("float", a[1])
            elif op == 23:
                self.types[a[0]] = ("vec", a[1], a[2])
            elif op == 24:
                self.types[a[0]] = ("mat", a[1], a[2])
            elif op == 25:
                self.types[a[0]] = ("image",)
            elif op == 26:
                self.types[a[0]] = ("sampler",)
            elif op == 27:
                self.types[a[0]] = ("image",)                      # sampled image: bound as one sampler object
            elif op == 29:
                self.types[a[0]] = ("rtarray", a[1])
            elif op == 5341:
                self.types[a[0]] = ("accel",)
            elif op == 28:
                self.types[a[0]] = ("array", a[1], a[2])
            elif op == 30:
                self.types[a[0]] = ("struct", list(a[1:]))
            elif op == 32:
                self.types[a[0]] = ("ptr", a[1], a[2])
            elif op == 33:
                self.types[a[0]] = ("fn",)
            elif op == 43:
                ty = self.types[a[0]]
                self.consts[a[1]] = f32(struct.unpack("<f", struct.pack("<I", a[2]))[0]) if ty[0] == "float" else (_s32(a[2]) if ty[2] else _u32(a[2]))
            elif op == 41:
                self.consts[a[1]] = True
            elif op == 42:
                self.consts[a[1]] = False
            elif op == 44:
                self.consts[a[1]] = [self.consts[c] for c in a[2:]]
            elif op == 46:
                self.consts[a[1]] = self.zero(a[0])
            elif op == 59 and cur is None:
                self.globals[a[1]] = (a[0], a[2])                  # result type (pointer), storage class
                if len(a) > 3:
                    self.global_init[a[1]] = a[3]
            elif op == 54:
                cur = {"id": a[1], "params": [], "blocks": {}, "order": [], "type": a[0]}
                self.functions[a[1]] = cur
                label = None
            elif op == 55:
                cur["params"].append(a[1])
            elif op == 248:
                label = a[0]
                cur["blocks"][label] = []
                cur["order"].append(label)
            elif op == 56:
                cur = None
            elif cur is not None:
                cur["blocks"][label].append((op, a))
            i += wc

    @staticmethod
    def _str(words):
        b = b"".join(struct.pack("<I", x) for x in words)
        return b.split(b"\0")[0].decode()

    def zero(self, ty):
        t = self.types[ty]
        if t[0] == "float":
            return f32(0)
        if t[0] == "int":
            return 0
        if t[0] == "bool":
            return False
        if t[0] == "vec":
            return [self.zero(t[1]) for _ in range(t[2])]
        if t[0] == "mat":
            return [self.zero(t[1]) for _ in range(t[2])]
        if t[0] == "array":
            return [self.zero(t[1]) for _ in range(self.consts[t[2]])]
        if t[0] == "struct":
            return [self.zero(m) for m in t[1]]
        if t[0] == "rtarray":
            return []
        if t[0] in ("image", "sampler", "accel"):
            return None
        raise NotImplementedError(t)

    def struct_member(self, struct_type_id, name):
        for (sid, k), n in self.member_names.items():
            if sid == struct_type_id and n == name:
                return k
        raise KeyError(name)


class Invocation:
    """Runs the entry point once.  `bind` maps global variable NAMES to values: images are objects with read(xy) -> [4 x f32]
    and write(xy, [4 x f32]); uniform / push-constant blocks are nested lists (Module.zero + struct_member help build them);
    gl_GlobalInvocationID is a 3-list of ints."""

    def __init__(self, module, bind):
        self.m = module
        self.g = {}
        self.trace = bind.get("traceRay")                                # callable(origin, direction, tmin, tmax, payload Ptr)
        for gid, (pty, storage) in module.globals.items():
            name = module.names.get(gid, "")
            pointee = module.types[pty][2]
            if name in bind:
                val = bind[name]
            elif any(d[0] == 11 and d[1] == 28 for d in module.decor.get(gid, [])):      # BuiltIn GlobalInvocationId
                val = bind["gl_GlobalInvocationID"]
            elif pointee in module.types and module.types[pointee][0] == "struct" and len(module.types[pointee][1]) == 1 and \
                    module.types[module.types[pointee][1][0]][0] == "rtarray":
                val = [[]]                                               # an unbound storage buffer: never indexed
            elif gid in module.global_init:
                val = _copy(module.consts[module.global_init[gid]])
            else:
                val = module.zero(pointee)
            self.g[gid] = Ptr([val], (0,))

    def run(self):
        self.call(self.m.entry, [])

    def call(self, fid, args):
        fn = self.m.functions[fid]
        v = dict(zip(fn["params"], args))
        m = self.m

        def val(i):
            if i in v:
                return v[i]
            if i in m.consts:
                return m.consts[i]
            if i in self.g:
                return self.g[i]
            raise KeyError(i)

        label, prev = fn["order"][0], None
        while True:
            nxt = None
            for op, a in fn["blocks"][label]:
                if op == 59:                                        # OpVariable (Function storage)
                    v[a[1]] = Ptr([m.zero(m.types[a[0]][2]) if len(a) < 4 else val(a[3])], (0,))
                elif op == 61:
                    x = val(a[2]).load()
                    v[a[1]] = list(x) if isinstance(x, list) else x
                elif op == 62:
                    x = val(a[1])
                    val(a[0]).store(_copy(x))
                elif op == 65:
                    v[a[1]] = val(a[2]).chain([int(val(k)) for k in a[3:]])
                elif op == 57:
                    v[a[1]] = self.call(a[2], [val(k) for k in a[3:]])
                elif op in (253, 254):
                    return val(a[0]) if op == 254 else None
                elif op in (246, 247):
                    pass
                elif op == 249:
                    nxt = a[0]
                elif op == 250:
                    nxt = a[1] if val(a[0]) else a[2]
                elif op == 245:                                     # OpPhi
                    pairs = a[2:]
                    v[a[1]] = next(val(pairs[k]) for k in range(0, len(pairs), 2) if pairs[k + 1] == prev)
                elif op == 79:
                    src = list(val(a[2])) + list(val(a[3]))
                    v[a[1]] = [src[k] for k in a[4:]]
                elif op == 80:
                    out = []
                    for k in a[2:]:
                        x = val(k)
                        out.extend(x) if isinstance(x, list) and m.types[a[0]][0] == "vec" else out.append(x)
                    v[a[1]] = out
                elif op == 81:
                    x = val(a[2])
                    for k in a[3:]:
                        x = x[k]
                    v[a[1]] = x
                elif op == 124:                                     # OpBitcast between 32-bit ints
                    signed = _is_signed(m, a[0])
                    v[a[1]] = _map(_s32 if signed else _u32, val(a[2]))
                elif op == 111:
                    v[a[1]] = _map(lambda x: f32(_s32(x)), val(a[2]))
                elif op == 112:
                    v[a[1]] = _map(lambda x: f32(_u32(x)), val(a[2]))
                elif op in (128, 130, 132):
                    fn2 = {128: lambda x, y: x + y, 130: lambda x, y: x - y, 132: lambda x, y: x * y}[op]
                    wrap = _s32 if _is_signed(m, a[0]) else _u32
                    v[a[1]] = _map(lambda x, y: wrap(fn2(int(x), int(y))), val(a[2]), val(a[3]))
                elif op == 135:                                     # OpSDiv: truncation toward zero
                    v[a[1]] = _map(lambda x, y: _s32(int(abs(_s32(x)) // abs(_s32(y))) * (1 if (_s32(x) < 0) == (_s32(y) < 0) else -1)), val(a[2]), val(a[3]))
                elif op in (129, 131, 133, 136):
                    fn2 = {129: lambda x, y: f32(x + y), 131: lambda x, y: f32(x - y), 133: lambda x, y: f32(x * y), 136: lambda x, y: f32(x / y)}[op]
                    v[a[1]] = _map(fn2, val(a[2]), val(a[3]))
                elif op == 127:
                    v[a[1]] = _map(lambda x: f32(-x), val(a[2]))
                elif op == 142:                                     # OpVectorTimesScalar
                    s = val(a[3])
                    v[a[1]] = [f32(x * s) for x in val(a[2])]
                elif op == 148:                                     # OpDot, summed left to right
                    x, y = val(a[2]), val(a[3])
                    acc = f32(x[0] * y[0])
                    for k in range(1, len(x)):
                        acc = f32(acc + f32(x[k] * y[k]))
                    v[a[1]] = acc
                elif op == 12:                                      # OpExtInst GLSL.std.450
                    inst, ops = a[3], [val(k) for k in a[4:]]
                    if inst == 4:
                        v[a[1]] = _map(lambda x: f32(abs(x)), ops[0])
                    elif inst == 40:
                        v[a[1]] = _map(lambda x, y: f32(max(x, y)), ops[0], ops[1])
                    elif inst == 69:
                        x = ops[0]
                        acc = f32(x[0] * x[0])
                        for k in range(1, len(x)):
                            acc = f32(acc + f32(x[k] * x[k]))
                        inv = f32(f32(1.0) / f32(np.sqrt(acc)))
                        v[a[1]] = [f32(c * inv) for c in x]
                    elif inst == 8:
                        v[a[1]] = _map(lambda x: f32(np.floor(x)), ops[0])
                    elif inst == 26:
                        v[a[1]] = _map(lambda x, y: f32(np.power(f32(x), f32(y))), ops[0], ops[1])
                    elif inst == 31:
                        v[a[1]] = _map(lambda x: f32(np.sqrt(f32(x))), ops[0])
                    elif inst == 37:
                        v[a[1]] = _map(lambda x, y: f32(min(x, y)), ops[0], ops[1])
                    elif inst == 43:                                # FClamp = min(max(x, lo), hi)
                        v[a[1]] = _map(lambda x, lo, hi: f32(min(max(x, lo), hi)), ops[0], ops[1], ops[2])
                    elif inst == 45:                                # SClamp
                        v[a[1]] = _map(lambda x, lo, hi: _s32(min(max(_s32(x), _s32(lo)), _s32(hi))), ops[0], ops[1], ops[2])
                    elif inst == 46:                                # FMix = x * (1 - a) + y * a
                        v[a[1]] = _map(lambda x, y, t: f32(f32(x * f32(f32(1.0) - t)) + f32(y * t)), ops[0], ops[1], ops[2])
                    elif inst == 67:                                # Distance
                        d = [f32(p - q) for p, q in zip(ops[0], ops[1])]
                        v[a[1]] = f32(np.sqrt(_dot(d, d)))
                    elif inst == 66:
                        v[a[1]] = f32(np.sqrt(_dot(ops[0], ops[0])))
                    elif inst == 71:                                # Reflect = I - 2 dot(N, I) N
                        dn = f32(f32(2.0) * _dot(ops[1], ops[0]))
                        v[a[1]] = [f32(i - f32(dn * n)) for i, n in zip(ops[0], ops[1])]
                    elif inst == 13:
                        v[a[1]] = _map(lambda x: f32(np.sin(f32(x))), ops[0])
                    elif inst == 14:
                        v[a[1]] = _map(lambda x: f32(np.cos(f32(x))), ops[0])
                    elif inst == 17:
                        v[a[1]] = _map(lambda x: f32(np.arccos(f32(x))), ops[0])
                    else:
                        raise NotImplementedError(f"GLSL.std.450 {inst}")
                elif op == 172:
                    v[a[1]] = _map(lambda x, y: _u32(x) > _u32(y), val(a[2]), val(a[3]))
                elif op == 174:
                    v[a[1]] = _map(lambda x, y: _u32(x) >= _u32(y), val(a[2]), val(a[3]))
                elif op == 176:
                    v[a[1]] = _map(lambda x, y: _u32(x) < _u32(y), val(a[2]), val(a[3]))
                elif op == 201:                                     # OpBitFieldInsert
                    base, ins, off, cnt = _u32(val(a[2])), _u32(val(a[3])), int(val(a[4])), int(val(a[5]))
                    mask = (((1 << cnt) - 1) << off) & 0xFFFFFFFF
                    r = (base & ~mask) | ((ins << off) & mask)
                    v[a[1]] = _s32(r) if _is_signed(m, a[0]) else _u32(r)
                elif op == 234:                                     # OpAtomicIAdd: invocations run one after another
                    ptr = val(a[2])
                    old_v = ptr.load()
                    ptr.store(_u32(int(old_v) + int(val(a[5]))))
                    v[a[1]] = old_v
                elif op == 177:
                    v[a[1]] = _map(lambda x, y: _s32(x) < _s32(y), val(a[2]), val(a[3]))
                elif op == 170:
                    v[a[1]] = _map(lambda x, y: _u32(x) == _u32(y), val(a[2]), val(a[3]))
                elif op == 186:
                    v[a[1]] = _map(lambda x, y: bool(x > y), val(a[2]), val(a[3]))
                elif op == 190:
                    v[a[1]] = _map(lambda x, y: bool(x >= y), val(a[2]), val(a[3]))
                elif op == 184:
                    v[a[1]] = _map(lambda x, y: bool(x < y), val(a[2]), val(a[3]))
                elif op == 168:
                    v[a[1]] = _map(lambda x: not x, val(a[2]))
                elif op == 166:
                    v[a[1]] = _map(lambda x, y: bool(x or y), val(a[2]), val(a[3]))
                elif op == 167:
                    v[a[1]] = _map(lambda x, y: bool(x and y), val(a[2]), val(a[3]))
                elif op == 169:
                    c, x, y = val(a[2]), val(a[3]), val(a[4])
                    v[a[1]] = [xx if cc else yy for cc, xx, yy in zip(c, x, y)] if isinstance(c, list) else (x if c else y)
                elif op == 199:
                    v[a[1]] = _map(lambda x, y: _u32(x) & _u32(y), val(a[2]), val(a[3]))
                elif op in (83, 400):                               # OpCopyObject / OpCopyLogical
                    v[a[1]] = _copy(val(a[2]))
                elif op == 88:                                      # OpImageSampleExplicitLod (Lod operand)
                    assert a[4] == 2, "only the Lod image operand is supported"
                    v[a[1]] = [f32(c) for c in val(a[2]).sample_lod([f32(c) for c in val(a[3])], f32(val(a[5])))]
                elif op == 100:                                     # OpImage
                    v[a[1]] = val(a[2])
                elif op == 103:                                     # OpImageQuerySizeLod
                    v[a[1]] = [int(c) for c in val(a[2]).size(int(val(a[3])))]
                elif op == 110:
                    v[a[1]] = _map(lambda x: _s32(int(np.trunc(x))), val(a[2]))
                elif op == 144:                                     # OpVectorTimesMatrix: result[c] = dot(v, column c)
                    vec, mat = val(a[2]), val(a[3])
                    v[a[1]] = [_dot(vec, col) for col in mat]
                elif op == 145:                                     # OpMatrixTimesVector: sum_c column c * v[c], left to right
                    mat, vec = val(a[2]), val(a[3])
                    out = [f32(mat[0][r] * vec[0]) for r in range(len(mat[0]))]
                    for c2 in range(1, len(mat)):
                        out = [f32(out[r] + f32(mat[c2][r] * vec[c2])) for r in range(len(out))]
                    v[a[1]] = out
                elif op == 146:                                     # OpMatrixTimesMatrix: column c of the result = left * (column c of right)
                    left, right = val(a[2]), val(a[3])
                    res = []
                    for col in right:
                        out = [f32(left[0][r] * col[0]) for r in range(len(left[0]))]
                        for k2 in range(1, len(left)):
                            out = [f32(out[r] + f32(left[k2][r] * col[k2])) for r in range(len(out))]
                        res.append(out)
                    v[a[1]] = res
                elif op == 154:
                    v[a[1]] = any(val(a[2]))
                elif op == 171:
                    v[a[1]] = _map(lambda x, y: _u32(x) != _u32(y), val(a[2]), val(a[3]))
                elif op == 173:
                    v[a[1]] = _map(lambda x, y: _s32(x) > _s32(y), val(a[2]), val(a[3]))
                elif op == 175:
                    v[a[1]] = _map(lambda x, y: _s32(x) >= _s32(y), val(a[2]), val(a[3]))
                elif op in (182, 183):
                    v[a[1]] = _map(lambda x, y: bool(x != y), val(a[2]), val(a[3]))
                elif op == 180:
                    v[a[1]] = _map(lambda x, y: bool(x == y), val(a[2]), val(a[3]))
                elif op == 188:
                    v[a[1]] = _map(lambda x, y: bool(x <= y), val(a[2]), val(a[3]))
                elif op in (202, 203):                              # OpBitFieldSExtract / UExtract
                    base, off, cnt = _u32(val(a[2])), int(val(a[3])), int(val(a[4]))
                    bits = (base >> off) & ((1 << cnt) - 1) if cnt else 0
                    if op == 202 and cnt and bits & (1 << (cnt - 1)):
                        bits -= 1 << cnt
                    v[a[1]] = _s32(bits) if op == 202 else _u32(bits)
                elif op == 4445:                                    # OpTraceRayKHR
                    self.trace([f32(c) for c in val(a[6])], [f32(c) for c in val(a[8])], f32(val(a[7])), f32(val(a[9])), val(a[10]))
                elif op == 98:                                      # OpImageRead
                    v[a[1]] = [f32(c) for c in val(a[2]).read([_s32(c) for c in val(a[3])])]
                elif op == 99:                                      # OpImageWrite
                    val(a[0]).write([_s32(c) for c in val(a[1])], [f32(c) for c in val(a[2])])
                else:
                    raise NotImplementedError(f"opcode {op}")
                if nxt is not None:
                    break
            if nxt is None:
                raise RuntimeError("block fell through")
            prev, label = label, nxt


def _dot(x, y):
    acc = f32(x[0] * y[0])
    for k in range(1, len(x)):
        acc = f32(acc + f32(x[k] * y[k]))
    return acc


def _is_signed(m, ty):
    t = m.types[ty]
    if t[0] == "vec":
        t = m.types[t[1]]
    return bool(t[2])


def _copy(x):
    return [_copy(c) for c in x] if isinstance(x, list) else x
