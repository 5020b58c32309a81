"""Binary op-table format shared by the Python lowering (``lowering.py``) and the CUDA
level interpreter (``csrc/rddl_plan.h`` mirrors every struct below field for field).

The op table is what the north star calls the "compact op table": the traced expression
DAG of every CPF level, the reward and the constraint checks of one RDDL problem, lowered
to a register bytecode plus the address descriptors, CSR neighbour lists and reduction
slots the kernels need. It is a flat little-endian blob:

    header  : int64[16]
    sections: tensors, launches, insns, consts, addrs, csrs, pool, redslots, variates
              (each padded to 8 bytes, in this order; counts are in the header)
"""
import numpy as np

MAGIC = 0x31304C444452  # "RDDL01"
VERSION = 11
MAX_IDX = 8          # index variables per thread: scope axes + nested loop axes
MAX_RED = 8          # reduction outputs per launch
MAX_REGS = 96        # virtual 64-bit registers per thread
MAX_TENSORS = 120

# --- opcodes ---------------------------------------------------------------
_OPS = [
    'NOP', 'CONST', 'AXIS', 'LOAD', 'STORE', 'MOV', 'I2F', 'F2I',
    # float
    'FADD', 'FSUB', 'FMUL', 'FDIV', 'FNEG', 'FABS', 'FMIN', 'FMAX', 'FPOW', 'FMOD', 'FLOGB',
    'FHYPOT', 'COS', 'SIN', 'TAN', 'ACOS', 'ASIN', 'ATAN', 'COSH', 'SINH', 'TANH', 'EXP', 'LN',
    'SQRT', 'LNGAMMA', 'GAMMA', 'FSGN', 'FROUND', 'FFLOOR', 'FCEIL',
    # int
    'IADD', 'ISUB', 'IMUL', 'INEG', 'IABS', 'IMIN', 'IMAX', 'IPOW', 'IFLOORDIV', 'IMOD', 'ISGN',
    # compare
    'FLT', 'FLE', 'FGT', 'FGE', 'FEQ', 'FNE', 'ILT', 'ILE', 'IGT', 'IGE', 'IEQ', 'INE',
    # logic on 0/1
    'AND', 'OR', 'XOR', 'NOT', 'IMPLY',
    'SEL',
    # control
    'LOOP', 'ENDLOOP', 'CSRLOOP', 'CSREND',
    # random base variates
    'RANDU', 'RANDN', 'RANDE', 'RANDC',
    # per-env reductions / status
    'REDOUT', 'REDIN', 'STATUS',
    'EXPM1', 'LOG1P',
    # bit-plane ops (rddl_plane_interp)
    'PLD', 'PLDC', 'PST', 'PCONST', 'PLUT', 'PSHARE', 'PCOUNT', 'PBERN', 'PRED',
    # scalar superinstruction: native multiply-accumulate loop (followed by one NOP data word)
    'SUMLOOP',
    # rejection / search samplers on a private Philox stream (flags = distribution id)
    'SAMPLE',
    # plane op the native library inserts for device-side policies with a max-nondef rule (never emitted here)
    'PPOLSEL', 'PPOL',
]
OP = {name: i for i, name in enumerate(_OPS)}

RED = {'SUM_I': 0, 'SUM_F': 1, 'PROD_I': 2, 'PROD_F': 3, 'MIN_I': 4, 'MIN_F': 5, 'MAX_I': 6,
       'MAX_F': 7, 'AND': 8, 'OR': 9}

DT = {'b': 0, 'i': 1, 'f': 2, 'packed': 3}      # 3: bool fluent stored as bits, 32 cells per uint32 word
KIND_SHARED, KIND_ENV = 0, 1

PLAN_STEP, PLAN_TERMINAL, PLAN_INVARIANT, PLAN_PRECONDITION = 0, 1, 2, 3

STATUS_OUT_OF_RANGE = 1      # simulator.py:229-253 -> RDDLValueOutOfRangeError
STATUS_BAD_BOUNDS = 2        # simulator.py:240-246
STATUS_BAD_PDF = 4           # simulator.py:1240-1252
STATUS_BAD_INDEX = 8         # nested object index outside its type (clamped)

INSN = np.dtype([('op', 'u1'), ('flags', 'u1'), ('dst', '<u2'), ('a', '<u2'), ('b', '<u2'),
                 ('c', '<u2'), ('pad', '<u2'), ('imm', '<u4')])
ADDR = np.dtype([('tensor', '<i4'), ('naxis', '<i4'), ('env_stride', '<i8'), ('c0', '<i8'),
                 ('axis', 'u1', (8,)), ('coef', '<i8', (8,)), ('nreg', '<i4'), ('pad', '<i4'),
                 ('reg', '<i4', (4,)), ('rstride', '<i8', (4,)), ('rextent', '<i8', (4,))])
CSR = np.dtype([('rowptr_off', '<i8'), ('ncols', '<i4'), ('col4_u16', '<i4'), ('col_off', '<i8', (4,)),
                ('slot', '<i4', (4,)), ('row_c0', '<i8'), ('row_naxis', '<i4'), ('pad2', '<i4'),
                ('row_axis', 'u1', (8,)), ('row_coef', '<i8', (8,)),
                # single-column lists: a second copy with every row padded to 16 bytes -- a multiple of four
                # int32 entries, or (col4_u16 = 1, when the indices fit) of eight uint16 entries; padding
                # entries = sentinel4, the extent of the reduced axis -- in the bank-scheduled order every
                # float SUMLOOP adds them in; rowptr4_off = 0: absent. rowptr4 counts entries.
                ('rowptr4_off', '<i8'), ('col4_off', '<i8'), ('sentinel4', '<i8')])
TENSOR = np.dtype([('dtype', '<i2'), ('pair', '<i2'), ('kind', '<i4'), ('nelem', '<i8')])  # pair: id of the primed partner of a state fluent, else -1
LAUNCH = np.dtype([('plan', '<i4'), ('rank', '<i4'), ('extent', '<i8', (8,)), ('E', '<i8'),
                   ('pc_begin', '<i4'), ('pc_end', '<i4'), ('nred', '<i4'), ('G', '<i4'),
                   ('ipt', '<i4'), ('chunks', '<i4'), ('red_op', '<i4', (8,)),
                   ('red_slot', '<i4', (8,)), ('nregs', '<i4'), ('pad', '<i4')])
REDSLOT = np.dtype([('op', '<i4'), ('chunks', '<i4')])
VARIATE = np.dtype([('node', '<i4'), ('kind', '<i4'), ('nelem', '<i8')])

assert INSN.itemsize == 16 and ADDR.itemsize == 184 and CSR.itemsize == 176
assert TENSOR.itemsize == 16 and LAUNCH.itemsize == 176 and REDSLOT.itemsize == 8


def pack(tensors, launches, insns, consts, addrs, csrs, pool, redslots, variates):
    """Serialises the sections into one blob (see module docstring)."""
    header = np.zeros(16, dtype='<i8')
    header[0] = MAGIC
    header[1] = VERSION
    secs = [np.ascontiguousarray(tensors, dtype=TENSOR), np.ascontiguousarray(launches, dtype=LAUNCH),
            np.ascontiguousarray(insns, dtype=INSN), np.ascontiguousarray(consts, dtype='<u8'),
            np.ascontiguousarray(addrs, dtype=ADDR), np.ascontiguousarray(csrs, dtype=CSR),
            np.ascontiguousarray(pool, dtype='<i4'), np.ascontiguousarray(redslots, dtype=REDSLOT),
            np.ascontiguousarray(variates, dtype=VARIATE)]
    for i, s in enumerate(secs):
        header[2 + i] = len(s)
    out = [header.tobytes()]
    for s in secs:
        b = s.tobytes()
        out.append(b)
        if len(b) % 8:
            out.append(b'\0' * (8 - len(b) % 8))
    return b''.join(out)

# distribution ids of the SAMPLE op (csrc/rddl_samplers.cuh: enum RddlDist)
SAMPLE_DIST = {'Poisson': 0, 'Gamma': 1, 'Binomial': 2, 'NegativeBinomial': 3, 'Beta': 4, 'Geometric': 5,
               'Student': 6, 'ChiSquare': 7}
