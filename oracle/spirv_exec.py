"""spirv_exec.py — a small SPIR-V 1.0 interpreter for the reference's shaders: the nine compute shaders of the path
and (Runner.run_stage) the vertex / fragment pair of its stage 5.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h). Purpose: execute the reference's own shipped
binaries — workdir/shaders/**/spv/*.comp.spv — on the CPU, because no Vulkan loader/ICD exists
here or on the GPU box. tools/make_spirv_golden.py drives it through the dispatch sequence of
ScanlineVGRasterizer::drawFrame (scanline_rasterizer.cpp:282-608) on small scenes and commits
every buffer as a golden fixture; tests/test_spirv_golden.py then pins the C oracle (and, on a
GPU, the CUDA path) to those buffers bit for bit.

tools/make_stage5_golden.py runs workdir/shaders/scanline/surface/spv/scanlinepr.{vert,frag}.spv per vertex / per
record over the reference's own output_buf dumps (tests/golden/stage5_*.npz, tests/test_stage5_golden.py).

Semantics implemented: the 84 opcodes glslang emitted for the nine compute shaders (+ OpImage / OpImageFetch on an integer
texel buffer, Input / Output variables and gl_PerVertex for the two graphics shaders) (logical addressing,
GLSL450 memory model), GLSL.std.450 {Floor, Sqrt, FAbs, FMin, FMax, SMin, SMax, SClamp}, std140 /
std430 buffer layouts from the Offset / ArrayStride decorations, Workgroup storage, barriers
(every invocation of a workgroup is a Python generator that yields at OpControlBarrier), and
OpAtomicExchange. Arithmetic: IEEE binary32 through numpy.float32, one rounding per instruction —
i.e. NO fused multiply-add, the "contract = none" policy of SURVEY App. D; OpDot sums left to right;
OpConvertFToS truncates and saturates. Invocations of a workgroup run in index order between
barriers, workgroups in dispatch order, so the reference's racy `path_visible[p] |= flag`
(transform_pos.comp:72-80) accumulates like an atomic OR.
"""
import struct

import numpy as np

# ---- opcodes (SPIR-V 1.0 unified spec) ---------------------------------------------------------
OP = dict(Name=5, MemberName=6, ExtInstImport=11, ExtInst=12, EntryPoint=15, ExecutionMode=16, TypeVoid=19, TypeBool=20,
          TypeInt=21, TypeFloat=22, TypeVector=23, TypeImage=25, TypeSampledImage=27, TypeArray=28, TypeRuntimeArray=29, TypeStruct=30, TypePointer=32,
          TypeFunction=33, ConstantTrue=41, ConstantFalse=42, Constant=43, ConstantComposite=44, Function=54,
          FunctionParameter=55, FunctionEnd=56, FunctionCall=57, Variable=59, Load=61, Store=62, AccessChain=65,
          Decorate=71, MemberDecorate=72, VectorShuffle=79, CompositeConstruct=80, CompositeExtract=81, ImageFetch=95, Image=100, ConvertFToU=109,
          ConvertFToS=110, ConvertSToF=111, ConvertUToF=112, Bitcast=124, SNegate=126, FNegate=127, IAdd=128, FAdd=129,
          ISub=130, FSub=131, IMul=132, FMul=133, UDiv=134, SDiv=135, FDiv=136, VectorTimesScalar=142, Dot=148,
          LogicalOr=166, LogicalAnd=167, LogicalNot=168, Select=169, IEqual=170, INotEqual=171, UGreaterThan=172,
          SGreaterThan=173, UGreaterThanEqual=174, SGreaterThanEqual=175, ULessThan=176, SLessThan=177, ULessThanEqual=178,
          SLessThanEqual=179, FOrdEqual=180, FUnordNotEqual=183, FOrdLessThan=184, FOrdGreaterThan=186,
          FOrdLessThanEqual=188, FOrdGreaterThanEqual=190, ShiftRightLogical=194, ShiftRightArithmetic=195,
          ShiftLeftLogical=196, BitwiseOr=197, BitwiseXor=198, BitwiseAnd=199, Not=200, ControlBarrier=224,
          MemoryBarrier=225, AtomicExchange=229, Phi=245, LoopMerge=246, SelectionMerge=247, Label=248, Branch=249,
          BranchConditional=250, Switch=251, Return=253, ReturnValue=254)
O = type("O", (), OP)
DEC_ARRAY_STRIDE, DEC_BUILTIN, DEC_BINDING, DEC_OFFSET = 6, 11, 33, 35
SC_UNIFORM_CONSTANT, SC_INPUT, SC_UNIFORM, SC_OUTPUT, SC_WORKGROUP, SC_PRIVATE, SC_FUNCTION, SC_PUSH, SC_STORAGE = 0, 1, 2, 3, 4, 6, 7, 9, 12
DEC_LOCATION = 30
BI_NUM_WG, BI_WG_ID, BI_LOCAL_ID, BI_GLOBAL_ID, BI_LOCAL_INDEX = 24, 26, 27, 28, 29
M32 = 0xFFFFFFFF
F32 = np.float32
_pk_f, _pk_i = struct.Struct("<f"), struct.Struct("<I")


def f2w(x):
    return _pk_i.unpack(_pk_f.pack(x))[0]


def w2f(w):
    return F32(_pk_f.unpack(_pk_i.pack(w & M32))[0])


def s32(u):
    u &= M32
    return u - (1 << 32) if u & 0x80000000 else u


def f2i_sat(x):  # truncation; out-of-range is undefined in SPIR-V: pinned like the oracle (saturate, NaN -> 0)
    x = float(x)
    if x != x:
        return 0
    if x >= 2147483648.0:
        return 2147483647
    if x <= -2147483648.0:
        return -2147483648
    return int(x)


class Type:
    __slots__ = ("kind", "width", "signed", "elem", "count", "members", "storage", "id")

    def __init__(self, kind, **kw):
        self.kind = kind
        self.width = self.signed = self.elem = self.count = self.members = self.storage = self.id = None
        for k, v in kw.items():
            setattr(self, k, v)


class Module:
    def __init__(self, path):
        data = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(data) // 4), data)
        assert w[0] == 0x07230203, "not SPIR-V"
        self.types, self.consts, self.names = {}, {}, {}
        self.decor, self.mdecor = {}, {}      # id -> {dec: [lits]};  (struct, member) -> {dec: [lits]}
        self.globals = {}                     # id -> (type_id, storage)
        self.functions = {}                   # id -> dict(params=[...], blocks={label: [instr]}, first=label, vars=[...])
        self.entry, self.local_size, self.glsl_ext = None, (1, 1, 1), None
        i, cur = 5, None
        cur_label = None
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            a = w[i + 1:i + wc]
            i += wc
            if op == O.Name:
                self.names[a[0]] = self._str(a[1:])
            elif op == O.ExtInstImport:
                self.glsl_ext = a[0]
            elif op == O.EntryPoint:
                self.entry = a[1]
            elif op == O.ExecutionMode:
                if a[1] == 17:
                    self.local_size = (a[2], a[3], a[4])
            elif op == O.Decorate:
                self.decor.setdefault(a[0], {})[a[1]] = list(a[2:])
            elif op == O.MemberDecorate:
                self.mdecor.setdefault((a[0], a[1]), {})[a[2]] = list(a[3:])
            elif op == O.TypeVoid:
                self.types[a[0]] = Type("void", id=a[0])
            elif op == O.TypeBool:
                self.types[a[0]] = Type("bool", id=a[0])
            elif op == O.TypeInt:
                self.types[a[0]] = Type("int", width=a[1], signed=bool(a[2]), id=a[0])
            elif op == O.TypeFloat:
                self.types[a[0]] = Type("float", width=a[1], id=a[0])
            elif op == O.TypeVector:
                self.types[a[0]] = Type("vector", elem=a[1], count=a[2], id=a[0])
            elif op == O.TypeImage:
                self.types[a[0]] = Type("image", elem=a[1], id=a[0])
            elif op == O.TypeSampledImage:
                self.types[a[0]] = Type("image", elem=a[1], id=a[0])
            elif op == O.TypeArray:
                self.types[a[0]] = Type("array", elem=a[1], count=("const", a[2]), id=a[0])
            elif op == O.TypeRuntimeArray:
                self.types[a[0]] = Type("array", elem=a[1], count=None, id=a[0])
            elif op == O.TypeStruct:
                self.types[a[0]] = Type("struct", members=list(a[1:]), id=a[0])
            elif op == O.TypePointer:
                self.types[a[0]] = Type("pointer", storage=a[1], elem=a[2], id=a[0])
            elif op == O.TypeFunction:
                self.types[a[0]] = Type("function", id=a[0])
            elif op in (O.ConstantTrue, O.ConstantFalse):
                self.consts[a[1]] = op == O.ConstantTrue
            elif op == O.Constant:
                t = self.types[a[0]]
                self.consts[a[1]] = w2f(a[2]) if t.kind == "float" else (a[2] & M32)
            elif op == O.ConstantComposite:
                self.consts[a[1]] = tuple(self.consts[x] for x in a[2:])
            elif op == O.Variable and cur is None:
                self.globals[a[1]] = (a[0], a[2])
            elif op == O.Function:
                cur = dict(params=[], blocks={}, first=None, vars=[], id=a[1])
                self.functions[a[1]] = cur
            elif op == O.FunctionParameter:
                cur["params"].append(a[1])
            elif op == O.FunctionEnd:
                cur = None
            elif cur is not None:
                if op == O.Label:
                    cur_label = a[0]
                    cur["blocks"][cur_label] = []
                    if cur["first"] is None:
                        cur["first"] = cur_label
                elif op == O.Variable:
                    cur["vars"].append((a[1], a[0]))
                elif op in (O.LoopMerge, O.SelectionMerge, O.MemoryBarrier):
                    pass
                else:
                    cur["blocks"][cur_label].append((op, a))
        for t in self.types.values():  # resolve array lengths
            if t.kind == "array" and t.count is not None:
                t.count = self.consts[t.count[1]]
        self._size_cache = {}

    @staticmethod
    def _str(words):
        b = b"".join(struct.pack("<I", x) for x in words)
        return b.split(b"\0", 1)[0].decode()

    # ---- layout (in 32-bit words) --------------------------------------------------------------
    def size_words(self, tid):
        if tid in self._size_cache:
            return self._size_cache[tid]
        t = self.types[tid]
        if t.kind in ("int", "float", "bool"):
            n = 1
        elif t.kind == "vector":
            n = t.count
        elif t.kind == "array":
            stride = self.decor.get(tid, {}).get(DEC_ARRAY_STRIDE)
            es = stride[0] // 4 if stride else self.size_words(t.elem)
            n = es * (t.count if t.count is not None else 0)
        elif t.kind == "struct":
            n = 0
            for m, mt in enumerate(t.members):
                off = self.mdecor.get((tid, m), {}).get(DEC_OFFSET)
                o = off[0] // 4 if off else n
                n = max(n, o + self.size_words(mt))
        else:
            raise ValueError(t.kind)
        self._size_cache[tid] = n
        return n

    def step(self, tid, index):
        """(word offset, element type) of component `index` inside a value of type tid."""
        t = self.types[tid]
        if t.kind == "vector":
            return index, t.elem
        if t.kind == "array":
            stride = self.decor.get(tid, {}).get(DEC_ARRAY_STRIDE)
            es = stride[0] // 4 if stride else self.size_words(t.elem)
            return index * es, t.elem
        if t.kind == "struct":
            off = self.mdecor.get((tid, index), {}).get(DEC_OFFSET)
            if off:
                return off[0] // 4, t.members[index]
            o = 0
            for m in range(index):
                o += self.size_words(t.members[m])
            return o, t.members[index]
        raise ValueError("cannot index " + t.kind)


class Runner:
    """Executes one dispatch of a compute module. `bindings[binding] = (uint32 ndarray, word_offset)`;
    `push` = uint32 ndarray for the push-constant block."""

    def __init__(self, module, bindings, push=None):
        self.m = module
        self.bind = {}
        self.global_ptr = {}
        self.instr_count = 0
        for vid, (ptid, sc) in module.globals.items():
            if sc in (SC_UNIFORM, SC_STORAGE):
                b = module.decor.get(vid, {}).get(DEC_BINDING)
                assert b is not None, "resource without binding"
                arr, off = bindings[b[0]]
                self.global_ptr[vid] = (arr, off, module.types[ptid].elem)
            elif sc == SC_PUSH:
                self.global_ptr[vid] = (push, 0, module.types[ptid].elem)
            elif sc == SC_UNIFORM_CONSTANT:  # a texel buffer (isamplerBuffer): bindings[b] = (int32 texels, 4 words each; 0)
                b = module.decor.get(vid, {}).get(DEC_BINDING)
                arr, off = bindings[b[0]]
                self.global_ptr[vid] = (arr, off, module.types[ptid].elem)

    # value <-> memory ---------------------------------------------------------------------------
    def load(self, ptr):
        mem, off, tid = ptr
        t = self.m.types[tid]
        k = t.kind
        if k == "float":
            return w2f(int(mem[off]))
        if k == "int":
            return int(mem[off]) & M32
        if k == "vector":
            et = self.m.types[t.elem].kind
            if et == "float":
                return tuple(w2f(int(mem[off + i])) for i in range(t.count))
            return tuple(int(mem[off + i]) & M32 for i in range(t.count))
        if k == "bool":
            return bool(mem[off])
        if k == "image":
            return ("image", mem, off)
        raise ValueError("load of " + k)

    def store(self, ptr, val):
        mem, off, tid = ptr
        t = self.m.types[tid]
        k = t.kind
        if k == "float":
            mem[off] = f2w(val)
        elif k == "int":
            mem[off] = val & M32
        elif k == "vector":
            if self.m.types[t.elem].kind == "float":
                for i in range(t.count):
                    mem[off + i] = f2w(val[i])
            else:
                for i in range(t.count):
                    mem[off + i] = val[i] & M32
        elif k == "bool":
            mem[off] = 1 if val else 0
        else:
            raise ValueError("store of " + k)

    # dispatch -----------------------------------------------------------------------------------
    def dispatch(self, gx, gy=1, gz=1):
        m = self.m
        lx, ly, lz = m.local_size
        n_local = lx * ly * lz
        for wz in range(gz):
            for wy in range(gy):
                for wx in range(gx):
                    shared = {}
                    for vid, (ptid, sc) in m.globals.items():
                        if sc == SC_WORKGROUP:
                            et = m.types[ptid].elem
                            shared[vid] = ([0] * m.size_words(et), 0, et)
                    gens = []
                    for li in range(n_local):
                        lid = (li % lx, (li // lx) % ly, li // (lx * ly))
                        env = dict(self.global_ptr)
                        env.update(shared)
                        for vid, (ptid, sc) in m.globals.items():
                            if sc == SC_INPUT:
                                bi = m.decor.get(vid, {}).get(DEC_BUILTIN, [None])[0]
                                val = {BI_WG_ID: (wx, wy, wz), BI_LOCAL_ID: lid, BI_NUM_WG: (gx, gy, gz),
                                       BI_GLOBAL_ID: (wx * lx + lid[0], wy * ly + lid[1], wz * lz + lid[2]),
                                       BI_LOCAL_INDEX: (li,)}[bi]
                                et = m.types[ptid].elem
                                env[vid] = (list(val), 0, et)
                            elif sc == SC_PRIVATE:
                                et = m.types[ptid].elem
                                env[vid] = ([0] * m.size_words(et), 0, et)
                        gens.append(self.call(m.entry, [], env))
                    live = gens
                    while live:  # run every invocation to its next barrier (or its end), in index order
                        nxt = []
                        for g in live:
                            try:
                                next(g)
                                nxt.append(g)
                            except StopIteration:
                                pass
                        live = nxt

    # one vertex / fragment invocation -----------------------------------------------------------
    def run_stage(self, builtins=None, locations=None):
        """Run the entry point once as a graphics-stage invocation (scanlinepr.vert / .frag: no barriers, no
        derivatives). Input variables are fed from `builtins` {BuiltIn number: tuple} (42 = VertexIndex, 15 =
        FragCoord) and `locations` {location: tuple} (floats as numpy.float32, integers as ints); returns
        ({BuiltIn: value}, {location: value}) read back from the Output variables (a gl_PerVertex block is reported
        member by member under its members' BuiltIn numbers, 0 = Position)."""
        m = self.m
        env = dict(self.global_ptr)
        outs = []
        for vid, (ptid, sc) in m.globals.items():
            if sc not in (SC_INPUT, SC_OUTPUT, SC_PRIVATE):
                continue
            et = m.types[ptid].elem
            mem = [0] * m.size_words(et)
            env[vid] = (mem, 0, et)
            dec = m.decor.get(vid, {})
            if sc == SC_INPUT:
                src = builtins.get(dec[DEC_BUILTIN][0]) if DEC_BUILTIN in dec else (locations or {}).get(dec.get(DEC_LOCATION, [None])[0])
                if src is not None:
                    self.store(env[vid], src if m.types[et].kind == "vector" else src[0])
            elif sc == SC_OUTPUT:
                outs.append((vid, et, dec))
        for _ in self.call(m.entry, [], env):
            raise RuntimeError("barrier in a graphics stage")
        out_b, out_l = {}, {}
        for vid, et, dec in outs:
            t = m.types[et]
            if t.kind == "struct":  # gl_PerVertex
                for k, mt in enumerate(t.members):
                    bi = m.mdecor.get((et, k), {}).get(DEC_BUILTIN)
                    if bi is not None and m.types[mt].kind in ("vector", "float", "int"):
                        o, _ = m.step(et, k)
                        out_b[bi[0]] = self.load((env[vid][0], o, mt))
            elif t.kind == "array":  # gl_SampleMask[]
                if DEC_BUILTIN in dec:
                    out_b[dec[DEC_BUILTIN][0]] = tuple(int(x) & M32 for x in env[vid][0])
            elif DEC_BUILTIN in dec:
                out_b[dec[DEC_BUILTIN][0]] = self.load(env[vid])
            elif DEC_LOCATION in dec:
                out_l[dec[DEC_LOCATION][0]] = self.load(env[vid])
        return out_b, out_l

    # interpreter --------------------------------------------------------------------------------
    def call(self, fid, args, genv):
        m = self.m
        fn = m.functions[fid]
        v = dict(genv)  # id -> value (pointers are (mem, off, type) tuples)
        consts = m.consts
        for pid, a in zip(fn["params"], args):
            v[pid] = a
        for vid, ptid in fn["vars"]:
            et = m.types[ptid].elem
            v[vid] = ([0] * m.size_words(et), 0, et)
        types = m.types

        def val(i):
            r = v.get(i)
            if r is None and i not in v:
                return consts[i]
            return r

        label, prev = fn["first"], None
        blocks = fn["blocks"]
        while True:
            block = blocks[label]
            nxt = None
            # phis first (they read values of the predecessor)
            pending = None
            for op, a in block:
                if op != O.Phi:
                    break
                for k in range(2, len(a), 2):
                    if a[k + 1] == prev:
                        if pending is None:
                            pending = []
                        pending.append((a[1], val(a[k])))
                        break
            if pending:
                for rid, x in pending:
                    v[rid] = x
            for op, a in block:
                self.instr_count += 1
                if op == O.Load:
                    v[a[1]] = self.load(val(a[2]))
                elif op == O.Store:
                    self.store(val(a[0]), val(a[1]))
                elif op == O.AccessChain:
                    mem, off, tid = val(a[2])
                    for ix in a[3:]:
                        o, tid = m.step(tid, val(ix))
                        off += o
                    v[a[1]] = (mem, off, tid)
                elif op == O.IAdd:
                    v[a[1]] = self._ew(val(a[2]), val(a[3]), lambda x, y: (x + y) & M32)
                elif op == O.ISub:
                    v[a[1]] = self._ew(val(a[2]), val(a[3]), lambda x, y: (x - y) & M32)
                elif op == O.IMul:
                    v[a[1]] = self._ew(val(a[2]), val(a[3]), lambda x, y: (x * y) & M32)
                elif op == O.FAdd:
                    v[a[1]] = self._ew(val(a[2]), val(a[3]), lambda x, y: x + y)
                elif op == O.FSub:
                    v[a[1]] = self._ew(val(a[2]), val(a[3]), lambda x, y: x - y)
                elif op == O.FMul:
                    v[a[1]] = self._ew(val(a[2]), val(a[3]), lambda x, y: x * y)
                elif op == O.FDiv:
                    with np.errstate(all="ignore"):
                        v[a[1]] = self._ew(val(a[2]), val(a[3]), lambda x, y: x / y)
                elif op == O.Phi:
                    pass
                elif op == O.Branch:
                    nxt = a[0]
                elif op == O.BranchConditional:
                    nxt = a[1] if val(a[0]) else a[2]
                elif op == O.Switch:
                    sel = val(a[0]) & M32
                    nxt = a[1]
                    for k in range(2, len(a), 2):
                        if a[k] == sel:
                            nxt = a[k + 1]
                            break
                elif op == O.Bitcast:
                    x = val(a[2])
                    tk = types[a[0]].kind
                    if tk == "float":
                        v[a[1]] = x if isinstance(x, F32) else w2f(x)
                    else:
                        v[a[1]] = f2w(x) if isinstance(x, F32) else (x & M32)
                elif op == O.BitwiseAnd:
                    v[a[1]] = val(a[2]) & val(a[3])
                elif op == O.BitwiseOr:
                    v[a[1]] = val(a[2]) | val(a[3])
                elif op == O.BitwiseXor:
                    v[a[1]] = val(a[2]) ^ val(a[3])
                elif op == O.Not:
                    v[a[1]] = (~val(a[2])) & M32
                elif op == O.ShiftLeftLogical:
                    v[a[1]] = (val(a[2]) << (val(a[3]) & 31)) & M32
                elif op == O.ShiftRightArithmetic:
                    v[a[1]] = (s32(val(a[2])) >> (val(a[3]) & 31)) & M32
                elif op == O.ShiftRightLogical:
                    v[a[1]] = (val(a[2]) & M32) >> (val(a[3]) & 31)
                elif op == O.IEqual:
                    v[a[1]] = (val(a[2]) & M32) == (val(a[3]) & M32)
                elif op == O.INotEqual:
                    v[a[1]] = (val(a[2]) & M32) != (val(a[3]) & M32)
                elif op == O.SLessThan:
                    v[a[1]] = s32(val(a[2])) < s32(val(a[3]))
                elif op == O.SLessThanEqual:
                    v[a[1]] = s32(val(a[2])) <= s32(val(a[3]))
                elif op == O.SGreaterThan:
                    v[a[1]] = s32(val(a[2])) > s32(val(a[3]))
                elif op == O.SGreaterThanEqual:
                    v[a[1]] = s32(val(a[2])) >= s32(val(a[3]))
                elif op == O.ULessThan:
                    v[a[1]] = (val(a[2]) & M32) < (val(a[3]) & M32)
                elif op == O.ULessThanEqual:
                    v[a[1]] = (val(a[2]) & M32) <= (val(a[3]) & M32)
                elif op == O.UGreaterThan:
                    v[a[1]] = (val(a[2]) & M32) > (val(a[3]) & M32)
                elif op == O.UGreaterThanEqual:
                    v[a[1]] = (val(a[2]) & M32) >= (val(a[3]) & M32)
                elif op == O.FOrdLessThan:
                    v[a[1]] = bool(val(a[2]) < val(a[3]))
                elif op == O.FOrdLessThanEqual:
                    v[a[1]] = bool(val(a[2]) <= val(a[3]))
                elif op == O.FOrdGreaterThan:
                    v[a[1]] = bool(val(a[2]) > val(a[3]))
                elif op == O.FOrdGreaterThanEqual:
                    v[a[1]] = bool(val(a[2]) >= val(a[3]))
                elif op == O.FOrdEqual:
                    v[a[1]] = self._cmp_eq(val(a[2]), val(a[3]))
                elif op == O.FUnordNotEqual:
                    x, y = val(a[2]), val(a[3])
                    v[a[1]] = tuple(not bool(p == q) for p, q in zip(x, y)) if isinstance(x, tuple) else (not bool(x == y))
                elif op == O.LogicalAnd:
                    v[a[1]] = bool(val(a[2])) and bool(val(a[3]))
                elif op == O.LogicalOr:
                    v[a[1]] = bool(val(a[2])) or bool(val(a[3]))
                elif op == O.LogicalNot:
                    v[a[1]] = not bool(val(a[2]))
                elif op == O.Select:
                    c, x, y = val(a[2]), val(a[3]), val(a[4])
                    v[a[1]] = tuple(p if k else q for k, p, q in zip(c, x, y)) if isinstance(c, tuple) else (x if c else y)
                elif op == O.ConvertFToS:
                    x = val(a[2])
                    v[a[1]] = tuple(f2i_sat(p) & M32 for p in x) if isinstance(x, tuple) else (f2i_sat(x) & M32)
                elif op == O.ConvertSToF:
                    v[a[1]] = F32(s32(val(a[2])))
                elif op == O.ConvertUToF:
                    v[a[1]] = F32(val(a[2]) & M32)
                elif op == O.FNegate:
                    x = val(a[2])
                    v[a[1]] = tuple(-p for p in x) if isinstance(x, tuple) else -x
                elif op == O.SNegate:
                    v[a[1]] = (-s32(val(a[2]))) & M32
                elif op == O.SDiv:
                    x, y = s32(val(a[2])), s32(val(a[3]))
                    q = abs(x) // abs(y) if y else 0
                    v[a[1]] = (q if (x < 0) == (y < 0) else -q) & M32
                elif op == O.UDiv:
                    v[a[1]] = ((val(a[2]) & M32) // (val(a[3]) & M32)) & M32
                elif op == O.CompositeConstruct:
                    parts = []
                    for x in a[2:]:
                        x = val(x)
                        parts.extend(x) if isinstance(x, tuple) else parts.append(x)
                    v[a[1]] = tuple(parts)
                elif op == O.CompositeExtract:
                    x = val(a[2])
                    for ix in a[3:]:
                        x = x[ix]
                    v[a[1]] = x
                elif op == O.Image:
                    v[a[1]] = val(a[2])
                elif op == O.ImageFetch:  # texelFetch on a buffer texture of 4-component 32-bit integer texels
                    _, mem, off = val(a[2])
                    i = s32(val(a[3]))
                    v[a[1]] = tuple(int(mem[off + 4 * i + k]) & M32 for k in range(4))
                elif op == O.VectorShuffle:
                    x, y = val(a[2]), val(a[3])
                    both = tuple(x) + tuple(y)
                    v[a[1]] = tuple(both[ix] for ix in a[4:])
                elif op == O.VectorTimesScalar:
                    x, s = val(a[2]), val(a[3])
                    v[a[1]] = tuple(p * s for p in x)
                elif op == O.Dot:
                    x, y = val(a[2]), val(a[3])
                    acc = x[0] * y[0]
                    for p, q in zip(x[1:], y[1:]):
                        acc = acc + p * q
                    v[a[1]] = acc
                elif op == O.ExtInst:
                    v[a[1]] = self._ext(a[3], [val(x) for x in a[4:]])
                elif op == O.FunctionCall:
                    r = yield from self.call(a[2], [val(x) for x in a[3:]], genv)
                    v[a[1]] = r
                elif op == O.ControlBarrier:
                    yield
                elif op == O.AtomicExchange:
                    ptr = val(a[2])
                    old = self.load(ptr)
                    self.store(ptr, val(a[5]))
                    v[a[1]] = old
                elif op == O.Return:
                    return None
                elif op == O.ReturnValue:
                    return val(a[0])
                else:
                    raise NotImplementedError("opcode %d" % op)
            prev, label = label, nxt

    @staticmethod
    def _ew(x, y, f):
        if isinstance(x, tuple):
            return tuple(f(p, q) for p, q in zip(x, y))
        return f(x, y)

    @staticmethod
    def _cmp_eq(x, y):
        if isinstance(x, tuple):
            return tuple(bool(p == q) for p, q in zip(x, y))
        return bool(x == y)

    def _ext(self, inst, a):
        def ew(f, *xs):
            if isinstance(xs[0], tuple):
                return tuple(f(*c) for c in zip(*xs))
            return f(*xs)
        if inst == 8:   # Floor
            return ew(lambda x: F32(np.floor(x)), a[0])
        if inst == 31:  # Sqrt
            with np.errstate(all="ignore"):
                return ew(lambda x: F32(np.sqrt(x)), a[0])
        if inst == 4:   # FAbs
            return ew(lambda x: F32(abs(x)), a[0])
        if inst == 37:  # FMin: y < x ? y : x
            return ew(lambda x, y: y if y < x else x, a[0], a[1])
        if inst == 40:  # FMax: x < y ? y : x
            return ew(lambda x, y: y if x < y else x, a[0], a[1])
        if inst == 39:  # SMin
            return ew(lambda x, y: (min(s32(x), s32(y))) & M32, a[0], a[1])
        if inst == 42:  # SMax
            return ew(lambda x, y: (max(s32(x), s32(y))) & M32, a[0], a[1])
        if inst == 45:  # SClamp
            return ew(lambda x, lo, hi: (min(max(s32(x), s32(lo)), s32(hi))) & M32, a[0], a[1], a[2])
        raise NotImplementedError("GLSL.std.450 instruction %d" % inst)
