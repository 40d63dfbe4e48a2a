"""A small SPIR-V interpreter -- TEST INFRASTRUCTURE, like everything under oracle/.

Why: the reference cannot be built or run here (MSVC project, un-vendored dependencies, needs a Vulkan ray-tracing
driver), but it ships the COMPILED shaders of the Hybrid path, `/root/reference/data/shaders/output/*.spv` (glslang,
SPIR-V 1.5, not optimised: every GLSL function is still an OpFunction).  Executing those binaries on the CPU gives outputs
of the reference itself: `tests/golden/make_spirv_vectors.py` runs individual functions (tea, lcg, rnd, sampleDisk, the GGX /
Smith / Schlick terms, cotangent_frame, perturbNormal) and whole entry points (blur.comp, postFragment.frag, miss.rmiss,
shadowRaygen.rgen, hybridRaygen.rgen, hybridHit.rchit, deferred.frag) on seeded inputs and commits the results as
`tests/golden/spirv_*.npz`; `tests/test_oracle_spirv.py` holds the scalar oracle (oracle/oracle.cpp) to them.  Nothing here is
read at test time on the GPU box: `/root/reference` does not exist there, only the committed vectors travel.

Semantics: 32-bit IEEE floats (numpy.float32, one rounding per operation, no contraction), 32-bit wrapping integers.
GLSL.std.450 transcendental functions are evaluated in float64 and rounded once (Vulkan leaves their precision to the
implementation; the oracle's own versions are compared with a tolerance).  Resources are supplied by the caller by name:
uniform / storage blocks as nested Python lists following the SPIR-V struct layout, images as objects with read / write /
sample methods, traceRayEXT as a callback.

Only what the reference's shaders use is implemented; an unknown opcode raises.
"""
import math
import struct

import numpy as np

F32, U32, I32 = np.float32, np.uint32, np.int32

BUILTINS = {0: "Position", 6: "InstanceId", 7: "PrimitiveId", 15: "FragCoord", 24: "NumWorkgroups", 25: "WorkgroupSize", 26: "WorkgroupId",
            27: "LocalInvocationId", 28: "GlobalInvocationId", 42: "VertexIndex", 43: "InstanceIndex",
            5319: "LaunchIdKHR", 5320: "LaunchSizeKHR", 5321: "WorldRayOriginKHR", 5322: "WorldRayDirectionKHR", 5323: "ObjectRayOriginKHR",
            5324: "ObjectRayDirectionKHR", 5325: "RayTminKHR", 5326: "RayTmaxKHR", 5327: "InstanceCustomIndexKHR", 5330: "ObjectToWorldKHR",
            5331: "WorldToObjectKHR", 5332: "HitTKHR", 5333: "HitKindKHR", 5351: "IncomingRayFlagsKHR"}


class Type:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)

    def __repr__(self):
        return f"Type({self.kind})"


class Ptr:
    """A pointer = a root cell (1-element list) plus a path of indices into nested lists / arrays."""
    __slots__ = ("cell", "path")

    def __init__(self, cell, path=()):
        self.cell, self.path = cell, tuple(path)

    def load(self):
        v = self.cell[0]
        for i in self.path:
            v = v[int(i)]
        return v

    def store(self, value):
        if not self.path:
            self.cell[0] = value
            return
        v = self.cell[0]
        for i in self.path[:-1]:
            v = v[int(i)]
        v[int(self.path[-1])] = value


class Return(Exception):
    def __init__(self, value=None):
        self.value = value


class Kill(Exception):
    pass


def _lit_string(ops):
    bs = b"".join(struct.pack("<I", x) for x in ops)
    s = bs.split(b"\0")[0]
    return s.decode(), len(s) // 4 + 1


def f32(x):
    return F32(x)


def _u(x):
    return int(x) & 0xffffffff


def _s(x):
    x = int(x) & 0xffffffff
    return x - (1 << 32) if x & 0x80000000 else x


def deep_copy(v):
    if isinstance(v, list):
        return [deep_copy(x) for x in v]
    if isinstance(v, np.ndarray):
        return v.copy()
    return v


class Module:
    def __init__(self, path):
        data = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(data) // 4), data)
        if w[0] != 0x07230203:
            raise ValueError("not a SPIR-V module")
        self.path = path
        self.version = w[1]
        self.names, self.member_names, self.decor, self.member_decor = {}, {}, {}, {}
        self.types, self.consts, self.globals, self.functions, self.entry_points = {}, {}, {}, {}, {}
        self.ext_sets = {}
        cur_fn = None
        cur_block = None
        i = 5
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xffff
            o = w[i + 1:i + wc]
            i += wc
            if cur_fn is not None:
                if op == 55:
                    cur_fn["params"].append((o[1], o[0]))
                elif op == 248:
                    cur_block = []
                    cur_fn["blocks"][o[0]] = cur_block
                    cur_fn["order"].append(o[0])
                elif op == 56:
                    cur_fn = None
                elif op in (8, 317):      # OpLine / OpNoLine
                    pass
                else:
                    cur_block.append((op, o))
                continue
            if op == 5:
                self.names[o[0]] = _lit_string(o[1:])[0]
            elif op == 6:
                self.member_names[(o[0], o[1])] = _lit_string(o[2:])[0]
            elif op == 11:
                self.ext_sets[o[0]] = _lit_string(o[1:])[0]
            elif op == 15:
                name, n = _lit_string(o[2:])
                self.entry_points[name] = (o[0], o[1], o[2 + n:])
            elif op == 71:
                self.decor.setdefault(o[0], {})[o[1]] = o[2:]
            elif op == 72:
                self.member_decor.setdefault((o[0], o[1]), {})[o[2]] = o[3:]
            elif op == 19:
                self.types[o[0]] = Type("void")
            elif op == 20:
                self.types[o[0]] = Type("bool")
            elif op == 21:
                self.types[o[0]] = Type("int", width=o[1], signed=bool(o[2]))
            elif op == 22:
                self.types[o[0]] = Type("float", width=o[1])
            elif op == 23:
                self.types[o[0]] = Type("vector", elem=self.types[o[1]], n=o[2])
            elif op == 24:
                self.types[o[0]] = Type("matrix", col=self.types[o[1]], n=o[2])
            elif op == 25:
                self.types[o[0]] = Type("image", dim=o[2], sampled=o[6])
            elif op == 26:
                self.types[o[0]] = Type("sampler")
            elif op == 27:
                self.types[o[0]] = Type("sampled_image")
            elif op == 28:
                self.types[o[0]] = Type("array", elem=self.types[o[1]], length_id=o[2])
            elif op == 29:
                self.types[o[0]] = Type("runtime_array", elem=self.types[o[1]])
            elif op == 30:
                self.types[o[0]] = Type("struct", members=[self.types[m] for m in o[1:]], id=o[0])
            elif op == 32:
                self.types[o[0]] = Type("pointer", storage=o[1], pointee=self.types[o[2]])
            elif op == 33:
                self.types[o[0]] = Type("function")
            elif op == 5341:
                self.types[o[0]] = Type("accel")
            elif op in (41, 42):
                self.consts[o[1]] = (op == 41)
            elif op == 43:
                t = self.types[o[0]]
                if t.kind == "float":
                    self.consts[o[1]] = F32(struct.unpack("<f", struct.pack("<I", o[2]))[0])
                elif t.signed:
                    self.consts[o[1]] = I32(_s(o[2]))
                else:
                    self.consts[o[1]] = U32(o[2])
            elif op == 44:
                self.consts[o[1]] = self._composite(self.types[o[0]], [self.consts[c] for c in o[2:]])
            elif op == 46:      # OpConstantNull
                self.consts[o[1]] = self.zero(self.types[o[0]])
            elif op == 59:
                self.globals[o[1]] = (self.types[o[0]], o[2], o[3] if len(o) > 3 else None)
            elif op == 54:
                cur_fn = {"id": o[1], "params": [], "blocks": {}, "order": [], "ret": self.types[o[0]]}
                self.functions[o[1]] = cur_fn
        self.fn_by_name = {}
        for fid in self.functions:
            n = self.names.get(fid, str(fid))
            self.fn_by_name[n.split("(")[0]] = fid

    # ---- values ----------------------------------------------------------------------------------------------
    def _composite(self, t, parts):
        if t.kind == "vector":
            dt = F32 if t.elem.kind == "float" else (np.bool_ if t.elem.kind == "bool" else (I32 if t.elem.signed else U32))
            flat = []
            for p in parts:
                if isinstance(p, np.ndarray):
                    flat.extend(p.tolist())
                else:
                    flat.append(p)
            return np.array(flat, dtype=dt)
        return [deep_copy(p) for p in parts]      # matrix = list of columns, struct / array = list

    def zero(self, t):
        k = t.kind
        if k == "float":
            return F32(0)
        if k == "int":
            return I32(0) if t.signed else U32(0)
        if k == "bool":
            return False
        if k == "vector":
            dt = F32 if t.elem.kind == "float" else (np.bool_ if t.elem.kind == "bool" else (I32 if t.elem.signed else U32))
            return np.zeros(t.n, dtype=dt)
        if k == "matrix":
            return [self.zero(t.col) for _ in range(t.n)]
        if k == "array":
            return [self.zero(t.elem) for _ in range(int(self.consts[t.length_id]))]
        if k == "struct":
            return [self.zero(m) for m in t.members]
        if k == "runtime_array":
            return []
        return None

    def array_length(self, t):
        return int(self.consts[t.length_id])

    def global_by_name(self, name):
        for gid, (t, storage, _init) in self.globals.items():
            if self.names.get(gid) == name:
                return gid
        # interface blocks are often anonymous instances of a named struct type
        for gid, (t, storage, _init) in self.globals.items():
            pt = t.pointee
            while pt.kind in ("array", "runtime_array"):
                pt = pt.elem
            if pt.kind == "struct" and self.names.get(pt.id) == name:
                return gid
        raise KeyError(name)

    def builtin_globals(self):
        out = {}
        for gid in self.globals:
            d = self.decor.get(gid, {})
            if 11 in d:
                out[BUILTINS.get(d[11][0], d[11][0])] = gid
        return out

    def describe_interface(self):
        """[(name, storage class, type kind, decorations)] of every global -- used to write the harnesses."""
        rows = []
        for gid, (t, storage, _i) in self.globals.items():
            pt = t.pointee
            tn = self.names.get(getattr(pt, "id", -1), pt.kind)
            rows.append((self.names.get(gid, ""), storage, tn, {k: v for k, v in self.decor.get(gid, {}).items()}))
        return rows


class Interpreter:
    """Executes functions of one module.  `resources`: {global variable name or block type name or builtin name: value}."""

    def __init__(self, module, resources=None, trace_ray=None, derivative=None):
        self.m = module
        self.trace_ray = trace_ray      # callable(interp, flags, cull, sbt_offset, sbt_stride, miss_index, origin, tmin, direction, tmax, payload_ptr)
        self.derivative = derivative    # callable(result_id, axis) -> value, for OpDPdx / OpDPdy
        self.cells = {}
        self.steps = 0
        m = module
        builtins = {gid: name for name, gid in m.builtin_globals().items()}
        resources = dict(resources or {})
        for gid, (t, storage, init) in m.globals.items():
            name = m.names.get(gid, "")
            val = None
            keys = [name, builtins.get(gid)]
            pt = t.pointee
            base = pt
            while base.kind in ("array", "runtime_array"):
                base = base.elem
            if base.kind == "struct":
                keys.append(m.names.get(base.id))
            found = False
            for k in keys:
                if k and k in resources:
                    val = resources[k]
                    found = True
                    break
            if not found:
                val = deep_copy(m.consts[init]) if init is not None else m.zero(pt)
            self.cells[gid] = [val]

    def global_value(self, name):
        return self.cells[self.m.global_by_name(name)][0]

    def set_global(self, name, value):
        self.cells[self.m.global_by_name(name)][0] = value

    # ---- running -------------------------------------------------------------------------------------------------
    def call(self, name, *args):
        """Call a function by its GLSL name with VALUES; returns (result, [final values of the pointer parameters])."""
        fid = self.m.fn_by_name[name]
        fn = self.m.functions[fid]
        ptrs = [Ptr([deep_copy(a)]) for a in args]
        r = self._run(fn, ptrs)
        return r, [p.load() for p in ptrs]

    def run_entry(self, name="main"):
        fid = self.m.entry_points[name][1]
        try:
            self._run(self.m.functions[fid], [])
        except Kill:
            return False
        return True

    def _run(self, fn, args):
        m = self.m
        v = {}
        for (pid, _t), a in zip(fn["params"], args):
            v[pid] = a

        def val(i):
            if i in v:
                return v[i]
            if i in m.consts:
                return m.consts[i]
            if i in self.cells:
                return Ptr(self.cells[i])
            raise KeyError(f"id {i} has no value")

        label = fn["order"][0]
        prev = None
        while True:
            block = fn["blocks"][label]
            nxt = None
            for op, o in block:
                self.steps += 1
                if op == 245:      # OpPhi
                    for k in range(2, len(o), 2):
                        if o[k + 1] == prev:
                            v[o[1]] = val(o[k])
                            break
                    else:
                        raise RuntimeError("phi without matching predecessor")
                elif op in (246, 247):      # merge declarations
                    pass
                elif op == 249:
                    nxt = o[0]
                elif op == 250:
                    nxt = o[1] if bool(val(o[0])) else o[2]
                elif op == 251:
                    sel = _s(val(o[0])) if isinstance(val(o[0]), (np.int32, int)) else _u(val(o[0]))
                    nxt = o[1]
                    for k in range(2, len(o), 2):
                        if _s(o[k]) == sel or o[k] == (sel & 0xffffffff):
                            nxt = o[k + 1]
                            break
                elif op == 253:
                    return None
                elif op == 254:
                    return val(o[0])
                elif op == 252:
                    raise Kill()
                elif op == 255:
                    raise RuntimeError("OpUnreachable executed")
                else:
                    self._exec(op, o, v, val)
            if nxt is None:
                raise RuntimeError("block without terminator")
            prev, label = label, nxt

    # ---- one instruction ---------------------------------------------------------------------------------------------
    def _exec(self, op, o, v, val):
        m = self.m
        T = m.types
        if op == 59:      # function-local OpVariable
            t = T[o[0]].pointee
            init = deep_copy(val(o[3])) if len(o) > 3 else m.zero(t)
            v[o[1]] = Ptr([init])
        elif op == 61:
            v[o[1]] = deep_copy(val(o[2]).load())
        elif op == 62:
            val(o[0]).store(deep_copy(val(o[1])))
        elif op == 63:      # OpCopyMemory
            val(o[0]).store(deep_copy(val(o[1]).load()))
        elif op in (65, 66):
            base = val(o[2])
            v[o[1]] = Ptr(base.cell, base.path + tuple(_s(val(i)) for i in o[3:]))
        elif op == 68:      # OpArrayLength
            s = val(o[2]).load()
            v[o[1]] = U32(len(s[o[3]]))
        elif op == 57:
            fn = m.functions[o[2]]
            v[o[1]] = self._run(fn, [val(a) for a in o[3:]])
        elif op == 12:
            v[o[1]] = self._ext(T[o[0]], o[3], [val(a) for a in o[4:]])
        elif op == 79:      # OpVectorShuffle
            a, b = val(o[2]), val(o[3])
            cat = np.concatenate([a, b])
            v[o[1]] = np.array([cat[k] if k != 0xffffffff else 0 for k in o[4:]], dtype=a.dtype)
        elif op == 80:
            v[o[1]] = m._composite(T[o[0]], [val(a) for a in o[2:]])
        elif op == 81:
            x = val(o[2])
            for k in o[3:]:
                x = x[k]
            v[o[1]] = deep_copy(x)
        elif op == 82:      # OpCompositeInsert
            obj = deep_copy(val(o[3]))
            tgt = obj
            for k in o[4:-1]:
                tgt = tgt[k]
            tgt[o[-1]] = deep_copy(val(o[2]))
            v[o[1]] = obj
        elif op in (83, 400):      # OpCopyObject / OpCopyLogical
            v[o[1]] = deep_copy(val(o[2]))
        elif op == 84:      # OpTranspose
            cols = val(o[2])
            n_rows = len(cols[0])
            v[o[1]] = [np.array([c[r] for c in cols], dtype=F32) for r in range(n_rows)]
        elif op in (87, 88):      # OpImageSample{Implicit,Explicit}Lod
            v[o[1]] = np.asarray(val(o[2]).sample(val(o[3])), dtype=F32)
        elif op == 98:
            v[o[1]] = np.asarray(val(o[2]).read(val(o[3])), dtype=F32)
        elif op == 99:
            val(o[0]).write(val(o[1]), val(o[2]))
        elif op == 86:      # OpSampledImage
            v[o[1]] = val(o[2])
        elif op in (207, 208):
            if self.derivative is None:
                raise RuntimeError("derivative requested but no provider given")
            v[o[1]] = self.derivative(o[2], 0 if op == 207 else 1, val(o[2]))
        elif op == 4445:      # OpTraceRayKHR
            a = [val(x) for x in o]
            self.trace_ray(self, _u(a[1]), _u(a[2]), _u(a[3]), _u(a[4]), _u(a[5]), a[6], a[7], a[8], a[9], a[10])
        elif op == 109:
            v[o[1]] = self._map1(val(o[2]), lambda x: U32(int(max(0.0, min(float(x), 4294967295.0))) if x == x else 0), U32)
        elif op == 110:
            v[o[1]] = self._map1(val(o[2]), lambda x: I32(int(max(-2147483648.0, min(float(x), 2147483647.0))) if x == x else 0), I32)
        elif op == 111:
            v[o[1]] = self._map1(val(o[2]), lambda x: F32(_s(x)), F32)
        elif op == 112:
            v[o[1]] = self._map1(val(o[2]), lambda x: F32(_u(x)), F32)
        elif op == 124:      # OpBitcast
            v[o[1]] = self._bitcast(T[o[0]], val(o[2]))
        elif op == 126:
            v[o[1]] = self._int1(T[o[0]], val(o[2]), lambda a: -a)
        elif op == 127:
            v[o[1]] = -val(o[2])
        elif op == 128:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: a + b)
        elif op == 130:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: a - b)
        elif op == 132:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: a * b)
        elif op == 134:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _u(a) // _u(b) if _u(b) else 0)
        elif op == 135:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: int(_s(a) / _s(b)) if _s(b) else 0)
        elif op == 137:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _u(a) % _u(b) if _u(b) else 0)
        elif op in (138, 139):
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: int(math.fmod(_s(a), _s(b))) if op == 138 else _s(a) % _s(b))
        elif op == 129:
            v[o[1]] = val(o[2]) + val(o[3])
        elif op == 131:
            v[o[1]] = val(o[2]) - val(o[3])
        elif op == 133:
            v[o[1]] = val(o[2]) * val(o[3])
        elif op == 136:
            with np.errstate(divide="ignore", invalid="ignore"):
                v[o[1]] = np.divide(val(o[2]), val(o[3]), dtype=F32)
        elif op == 141:      # OpFMod
            a, b = val(o[2]), val(o[3])
            v[o[1]] = a - b * np.floor(a / b)
        elif op == 142:
            v[o[1]] = val(o[2]) * val(o[3])
        elif op == 143:
            s = val(o[3])
            v[o[1]] = [c * s for c in val(o[2])]
        elif op == 145:      # matrix * vector: sum of columns scaled, in GLSL order (left to right)
            M, x = val(o[2]), val(o[3])
            acc = M[0] * x[0]
            for k in range(1, len(M)):
                acc = acc + M[k] * x[k]
            v[o[1]] = acc
        elif op == 144:      # vector * matrix
            x, M = val(o[2]), val(o[3])
            v[o[1]] = np.array([self._dot(x, c) for c in M], dtype=F32)
        elif op == 146:      # matrix * matrix
            A, B = val(o[2]), val(o[3])
            cols = []
            for bc in B:
                acc = A[0] * bc[0]
                for k in range(1, len(A)):
                    acc = acc + A[k] * bc[k]
                cols.append(acc)
            v[o[1]] = cols
        elif op == 148:
            v[o[1]] = self._dot(val(o[2]), val(o[3]))
        elif op in (154, 155):
            v[o[1]] = bool(np.any(val(o[2]))) if op == 154 else bool(np.all(val(o[2])))
        elif op == 156:
            v[o[1]] = self._cmp(val(o[2]), val(o[2]), lambda a, b: a != b)
        elif op == 157:
            v[o[1]] = self._map1(val(o[2]), lambda a: bool(np.isinf(a)), np.bool_)
        elif op == 164:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: bool(a) == bool(b))
        elif op == 165:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: bool(a) != bool(b))
        elif op == 166:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: bool(a) or bool(b))
        elif op == 167:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: bool(a) and bool(b))
        elif op == 168:
            v[o[1]] = self._map1(val(o[2]), lambda a: not bool(a), np.bool_)
        elif op == 169:
            c, a, b = val(o[2]), val(o[3]), val(o[4])
            if isinstance(c, np.ndarray):
                v[o[1]] = np.where(c, a, b).astype(a.dtype)
            else:
                v[o[1]] = deep_copy(a if bool(c) else b)
        elif op == 170:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _u(a) == _u(b))
        elif op == 171:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _u(a) != _u(b))
        elif op == 172:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _u(a) > _u(b))
        elif op == 173:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _s(a) > _s(b))
        elif op == 174:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _u(a) >= _u(b))
        elif op == 175:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _s(a) >= _s(b))
        elif op == 176:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _u(a) < _u(b))
        elif op == 177:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _s(a) < _s(b))
        elif op == 178:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _u(a) <= _u(b))
        elif op == 179:
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: _s(a) <= _s(b))
        elif op in (180, 181):
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: a == b)
        elif op in (182, 183):
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: a != b)
        elif op in (184, 185):
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: a < b)
        elif op in (186, 187):
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: a > b)
        elif op in (188, 189):
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: a <= b)
        elif op in (190, 191):
            v[o[1]] = self._cmp(val(o[2]), val(o[3]), lambda a, b: a >= b)
        elif op == 194:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _u(a) >> (_u(b) & 31))
        elif op == 195:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _s(a) >> (_u(b) & 31))
        elif op == 196:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _u(a) << (_u(b) & 31))
        elif op == 197:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _u(a) | _u(b))
        elif op == 198:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _u(a) ^ _u(b))
        elif op == 199:
            v[o[1]] = self._int2(T[o[0]], val(o[2]), val(o[3]), lambda a, b: _u(a) & _u(b))
        elif op == 200:
            v[o[1]] = self._int1(T[o[0]], val(o[2]), lambda a: ~_u(a))
        else:
            raise NotImplementedError(f"SPIR-V opcode {op}")

    # ---- helpers -----------------------------------------------------------------------------------------------
    @staticmethod
    def _dot(a, b):
        p = a * b
        acc = p[0]
        for k in range(1, len(p)):
            acc = acc + p[k]
        return F32(acc)

    @staticmethod
    def _map1(a, fn, dt):
        if isinstance(a, np.ndarray):
            return np.array([fn(x) for x in a], dtype=dt)
        return fn(a)

    @staticmethod
    def _cmp(a, b, fn):
        if isinstance(a, np.ndarray):
            bb = b if isinstance(b, np.ndarray) else [b] * len(a)
            return np.array([bool(fn(x, y)) for x, y in zip(a, bb)], dtype=np.bool_)
        return bool(fn(a, b))

    @staticmethod
    def _wrap(t, x):
        e = t.elem if t.kind == "vector" else t
        return I32(_s(x)) if e.signed else U32(_u(x))

    def _int2(self, t, a, b, fn):
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            n = len(a) if isinstance(a, np.ndarray) else len(b)
            aa = a if isinstance(a, np.ndarray) else [a] * n
            bb = b if isinstance(b, np.ndarray) else [b] * n
            e = t.elem
            return np.array([self._wrap(t, fn(int(x), int(y))) for x, y in zip(aa, bb)], dtype=I32 if e.signed else U32)
        return self._wrap(t, fn(int(a), int(b)))

    def _int1(self, t, a, fn):
        if isinstance(a, np.ndarray):
            return np.array([self._wrap(t, fn(int(x))) for x in a], dtype=I32 if t.elem.signed else U32)
        return self._wrap(t, fn(int(a)))

    @staticmethod
    def _bitcast(t, x):
        e = t.elem if t.kind == "vector" else t
        dt = F32 if e.kind == "float" else (I32 if e.signed else U32)
        if isinstance(x, np.ndarray):
            return x.view(dt).copy()
        return np.array([x]).view(dt)[0]

    def _ext(self, t, inst, a):
        def per(fn, *xs):
            if isinstance(xs[0], np.ndarray):
                n = len(xs[0])
                ys = [x if isinstance(x, np.ndarray) else [x] * n for x in xs]
                return np.array([fn(*[y[k] for y in ys]) for k in range(n)], dtype=xs[0].dtype)
            return fn(*xs)

        def d(fn):      # float64 evaluation, one rounding
            def g(*xs):
                try:
                    return F32(fn(*[float(x) for x in xs]))
                except (ValueError, ZeroDivisionError, OverflowError):
                    return F32(np.nan)
            return g

        if inst == 1:
            return per(lambda x: F32(math.floor(float(x) + 0.5)), a[0])
        if inst == 2:
            return per(lambda x: F32(np.round(x)), a[0])
        if inst == 3:
            return per(lambda x: F32(np.trunc(x)), a[0])
        if inst == 4:
            return np.abs(a[0])
        if inst == 5:
            return per(lambda x: I32(abs(_s(x))), a[0])
        if inst == 6:
            return per(lambda x: F32(np.sign(x)), a[0])
        if inst == 8:
            return np.floor(a[0])
        if inst == 9:
            return np.ceil(a[0])
        if inst == 10:
            return a[0] - np.floor(a[0])
        if inst == 11:
            return a[0] * F32(0.017453292519943295)
        if inst == 12:
            return a[0] * F32(57.29577951308232)
        if inst == 13:
            return per(d(math.sin), a[0])
        if inst == 14:
            return per(d(math.cos), a[0])
        if inst == 15:
            return per(d(math.tan), a[0])
        if inst == 16:
            return per(d(math.asin), a[0])
        if inst == 17:
            return per(d(math.acos), a[0])
        if inst == 18:
            return per(d(math.atan), a[0])
        if inst == 25:
            return per(d(math.atan2), a[0], a[1])
        if inst == 26:
            return per(d(lambda x, y: math.pow(x, y) if not (x < 0) else float("nan")), a[0], a[1])
        if inst == 27:
            return per(d(math.exp), a[0])
        if inst == 28:
            return per(d(lambda x: math.log(x) if x > 0 else (float("-inf") if x == 0 else float("nan"))), a[0])
        if inst == 29:
            return per(d(lambda x: 2.0 ** x), a[0])
        if inst == 30:
            return per(d(lambda x: math.log2(x) if x > 0 else (float("-inf") if x == 0 else float("nan"))), a[0])
        if inst == 31:
            return np.sqrt(a[0])
        if inst == 32:
            with np.errstate(divide="ignore", invalid="ignore"):
                return np.divide(F32(1), np.sqrt(a[0]), dtype=F32)
        if inst == 33:      # Determinant
            return F32(np.linalg.det(np.array([c.astype(np.float64) for c in a[0]]).T))
        if inst == 34:      # MatrixInverse: float64, rounded once (precision is implementation-defined)
            M = np.array([c.astype(np.float64) for c in a[0]]).T
            inv = np.linalg.inv(M)
            return [inv[:, k].astype(F32) for k in range(inv.shape[1])]
        if inst == 37:
            return np.where(a[1] < a[0], a[1], a[0]).astype(F32) if isinstance(a[0], np.ndarray) else (a[1] if a[1] < a[0] else a[0])
        if inst == 40:
            return np.where(a[0] < a[1], a[1], a[0]).astype(F32) if isinstance(a[0], np.ndarray) else (a[1] if a[0] < a[1] else a[0])
        if inst in (38, 39):
            conv = _u if inst == 38 else _s
            return per(lambda x, y: x if conv(x) <= conv(y) else y, a[0], a[1])
        if inst in (41, 42):
            conv = _u if inst == 41 else _s
            return per(lambda x, y: x if conv(x) >= conv(y) else y, a[0], a[1])
        if inst == 43:      # FClamp = min(max(x, lo), hi)
            return per(lambda x, lo, hi: min(max(x, lo), hi), a[0], a[1], a[2])
        if inst in (44, 45):
            conv = _u if inst == 44 else _s
            return per(lambda x, lo, hi: min(max(x, lo, key=conv), hi, key=conv), a[0], a[1], a[2])
        if inst == 46:      # FMix = x * (1 - a) + y * a
            return a[0] * (F32(1) - a[2]) + a[1] * a[2]
        if inst == 48:
            return per(lambda e, x: F32(0) if x < e else F32(1), a[0], a[1])
        if inst == 49:
            def ss(e0, e1, x):
                tt = min(max((x - e0) / (e1 - e0), F32(0)), F32(1))
                return tt * tt * (F32(3) - F32(2) * tt)
            return per(ss, a[0], a[1], a[2])
        if inst == 50:
            return per(lambda x, y, z: F32(float(x) * float(y) + float(z)), a[0], a[1], a[2])
        if inst == 66:
            x = a[0]
            return F32(np.sqrt(self._dot(x, x))) if isinstance(x, np.ndarray) else np.abs(x)
        if inst == 67:
            x = a[0] - a[1]
            return F32(np.sqrt(self._dot(x, x)))
        if inst == 68:
            x, y = a
            return np.array([x[1] * y[2] - y[1] * x[2], x[2] * y[0] - y[2] * x[0], x[0] * y[1] - y[0] * x[1]], dtype=F32)
        if inst == 69:      # normalize = x / length(x), computed as x * inversesqrt(dot(x, x))
            x = a[0]
            with np.errstate(divide="ignore", invalid="ignore"):
                return x * np.divide(F32(1), np.sqrt(self._dot(x, x)), dtype=F32)
        if inst == 71:      # reflect: I - 2 * dot(N, I) * N
            I, N = a
            return I - F32(2) * self._dot(N, I) * N
        if inst == 72:      # refract
            I, N, eta = a
            dNI = self._dot(N, I)
            k = F32(1) - eta * eta * (F32(1) - dNI * dNI)
            if k < 0:
                return np.zeros_like(I)
            return eta * I - (eta * dNI + np.sqrt(k)) * N
        raise NotImplementedError(f"GLSL.std.450 instruction {inst}")


# ---- resource stubs -------------------------------------------------------------------------------------------------
class Image2D:
    """storage image / sampled texture over a float32 H x W x 4 array.  NEAREST filtering; out-of-range imageLoad gives 0,
    out-of-range imageStore is dropped (robust buffer access); texture() wraps with REPEAT unless clamp=True."""

    def __init__(self, array, unorm8=False, repeat=True, f16=False, corner_eps=0.0, write_to=None):
        self.a = np.asarray(array, dtype=F32).copy()
        self.unorm8, self.repeat, self.f16 = unorm8, repeat, f16
        self.corner_eps = corner_eps      # texture() exactly on a texel corner (uv = x / W) is implementation-defined: > 0 snaps up
        self.out = write_to               # imageStore target when reads must keep seeing the previous contents (ping-pong)
        self.reads, self.writes = [], []

    def read(self, coord):
        x, y = int(_s(coord[0])), int(_s(coord[1]))
        self.reads.append((x, y))
        h, w = self.a.shape[:2]
        if x < 0 or y < 0 or x >= w or y >= h:
            return np.zeros(4, F32)
        return self.a[y, x].copy()

    def write(self, coord, texel):
        x, y = int(_s(coord[0])), int(_s(coord[1]))
        self.writes.append((x, y))
        h, w = self.a.shape[:2]
        if x < 0 or y < 0 or x >= w or y >= h:
            return
        t = np.asarray(texel, dtype=F32)
        if self.unorm8:
            t = (np.rint(np.clip(t, 0, 1) * F32(255)) / F32(255)).astype(F32)
        if self.f16:
            t = t.astype(np.float16).astype(F32)
        (self.a if self.out is None else self.out)[y, x] = t

    def sample(self, uv):
        h, w = self.a.shape[:2]
        fx, fy = F32(uv[0]) * F32(w), F32(uv[1]) * F32(h)
        ix = int(math.floor(float(fx) + self.corner_eps)) if fx == fx else 0
        iy = int(math.floor(float(fy) + self.corner_eps)) if fy == fy else 0
        if self.repeat:
            ix %= w
            iy %= h
        else:
            ix = min(max(ix, 0), w - 1)
            iy = min(max(iy, 0), h - 1)
        return self.a[iy, ix].copy()
