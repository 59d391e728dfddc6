"""R1CS synthesis restated for the reference's test/bench circuits (oracle; test infrastructure).

A minimal stand-in for ark-relations 0.4 `ConstraintSystem` (un-vendored
dependency) as driven by `/root/reference/src/prover.rs:33-52` and
`src/generator.rs:31-54`: variables are allocated in call order, linear
combinations are kept sorted by variable (One < Instance(i) < Witness(i)) with
duplicates merged, `to_matrices` drops zero coefficients and maps One->0,
Instance(i)->i, Witness(i)->m0+i.

Circuits: `DummyCircuit` (tests/dummy.rs:20-35), `BenchDummyCircuit`
(benches/bench.rs:19-61) and `MiMCDemo` (tests/mimc.rs:46-143).
"""
from .fields import R_MOD

ONE = ("one", 0)


def inst(i):
    return ("inst", i)


def wit(i):
    return ("wit", i)


_KIND_ORDER = {"one": 0, "inst": 1, "wit": 2}


class ConstraintSystem:
    def __init__(self, setup_mode: bool):
        self.setup_mode = setup_mode
        self.instance_assignment = [1]
        self.witness_assignment = []
        self.num_instance_variables = 1
        self.num_witness_variables = 0
        self.constraints = []  # (a_lc, b_lc, c_lc), each a dict var->coeff
        self.num_constraints = 0

    def new_witness_variable(self, value_fn):
        idx = self.num_witness_variables
        self.num_witness_variables += 1
        if not self.setup_mode:
            self.witness_assignment.append(value_fn() % R_MOD)
        return wit(idx)

    def new_input_variable(self, value_fn):
        idx = self.num_instance_variables
        self.num_instance_variables += 1
        if not self.setup_mode:
            self.instance_assignment.append(value_fn() % R_MOD)
        return inst(idx)

    def enforce_constraint(self, a, b, c):
        """a, b, c: lists of (coeff, var)."""
        self.num_constraints += 1
        if self.setup_mode:
            self.constraints.append(tuple(self._lc(t) for t in (a, b, c)))

    @staticmethod
    def _lc(terms):
        d = {}
        for coeff, var in terms:
            d[var] = (d.get(var, 0) + coeff) % R_MOD
        return d

    def to_matrices(self):
        m0 = self.num_instance_variables

        def col(var):
            kind, i = var
            return 0 if kind == "one" else (i if kind == "inst" else m0 + i)

        def row(lc):
            items = sorted(lc.items(), key=lambda kv: (_KIND_ORDER[kv[0][0]], kv[0][1]))
            return [(c, col(v)) for v, c in items if c != 0]

        a = [row(t[0]) for t in self.constraints]
        b = [row(t[1]) for t in self.constraints]
        c = [row(t[2]) for t in self.constraints]
        return a, b, c


class DummyCircuit:
    """tests/dummy.rs:20-35 — one constraint a*b = c (c public)."""

    def __init__(self, a=None, b=None):
        self.a, self.b = a, b

    def generate_constraints(self, cs):
        a = cs.new_witness_variable(lambda: self.a)
        b = cs.new_witness_variable(lambda: self.b)
        c = cs.new_input_variable(lambda: self.a * self.b)
        cs.enforce_constraint([(1, a)], [(1, b)], [(1, c)])


class BenchDummyCircuit:
    """benches/bench.rs:19-61 — a*b=c repeated, unused witness copies of a, one empty row."""

    def __init__(self, a, b, num_variables, num_constraints):
        self.a, self.b = a, b
        self.num_variables, self.num_constraints = num_variables, num_constraints

    def generate_constraints(self, cs):
        a = cs.new_witness_variable(lambda: self.a)
        b = cs.new_witness_variable(lambda: self.b)
        c = cs.new_input_variable(lambda: self.a * self.b)
        for _ in range(self.num_variables - 3):
            cs.new_witness_variable(lambda: self.a)
        for _ in range(self.num_constraints - 1):
            cs.enforce_constraint([(1, a)], [(1, b)], [(1, c)])
        cs.enforce_constraint([], [], [])


def mimc_hash(xl, xr, constants):
    """tests/mimc.rs:46-62."""
    for c in constants:
        t = (xl + c) % R_MOD
        xl, xr = (t * t % R_MOD * t + xr) % R_MOD, xl
    return xl


class MiMCDemo:
    """tests/mimc.rs:66-143 with a configurable round count (322 in the reference)."""

    def __init__(self, xl, xr, constants):
        self.xl, self.xr, self.constants = xl, xr, constants

    def generate_constraints(self, cs):
        have = self.xl is not None
        xl_value, xr_value = self.xl, self.xr
        xl = cs.new_witness_variable(lambda: xl_value)
        xr = cs.new_witness_variable(lambda: xr_value)
        rounds = len(self.constants)
        for i, ci in enumerate(self.constants):
            tmp_value = pow(xl_value + ci, 2, R_MOD) if have else None
            tmp = cs.new_witness_variable(lambda: tmp_value)
            cs.enforce_constraint([(1, xl), (ci, ONE)], [(1, xl), (ci, ONE)], [(1, tmp)])
            new_xl_value = ((xl_value + ci) * tmp_value + xr_value) % R_MOD if have else None
            if i == rounds - 1:
                new_xl = cs.new_input_variable(lambda: new_xl_value)
            else:
                new_xl = cs.new_witness_variable(lambda: new_xl_value)
            cs.enforce_constraint([(1, tmp)], [(1, xl), (ci, ONE)], [(1, new_xl), (R_MOD - 1, xr)])
            xr, xr_value = xl, xl_value
            xl, xl_value = new_xl, new_xl_value


def synthesize(circuit, setup_mode: bool) -> ConstraintSystem:
    cs = ConstraintSystem(setup_mode)
    circuit.generate_constraints(cs)
    return cs
