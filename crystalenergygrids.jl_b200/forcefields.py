"""Pair-interaction model and its flattening into the device pair table.

Mirrors /root/reference/src/interactions.jl and forcefields.jl:
  FF.InteractionKind (enum order matters: rules in a sum are sorted by it) interactions.jl:23-33
  InteractionRule / rule(r)                          interactions.jl:232-237, 367-387
  _shifted_InteractionRule                           interactions.jl:350-365
  InteractionRuleSum                                 interactions.jl:558-582
  _mix_rules / reduce_rule_sum! / mix_rules          forcefields.jl:51-185
  ForceField ctor                                    forcefields.jl:204-265
  needsvdwgrid                                       forcefields.jl:306-316

The per-pair energy/derivative formulas themselves (derivativesGrid,
interactions.jl:432-472) are evaluated only on the device (csrc/) and in oracle/.
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field
from typing import Union

import numpy as np

from .constants import COULOMBIC_CONVERSION_FACTOR


class Kind(enum.IntEnum):
    HardSphere = 0
    CoulombEwaldDirect = 1
    Coulomb = 2
    LennardJones = 3
    Buckingham = 4
    Monomial = 5
    Exponential = 6
    UndefinedInteraction = 7
    NoInteraction = 8


class Mixing(enum.Enum):
    LorentzBerthelot = 1
    WaldmanHagler = 2
    Geometric = 3
    IgnoreInteraction = 4
    ErrorOnMix = 5


class UndefinedInteractionError(Exception):
    def __str__(self):
        return "Undefined interaction"


@dataclass(frozen=True)
class InteractionRule:
    kind: Kind
    params: tuple = ()
    shift: float = 0.0
    tailcorrection: bool = True

    def __call__(self, r: float) -> float:
        """interactions.jl:367-387 — energy in K at distance r (Å), minus the shift."""
        k, p = self.kind, self.params
        if k is Kind.LennardJones:
            x6 = (p[1] / r) ** 6
            v = 4 * p[0] * x6 * (x6 - 1)
        elif k is Kind.CoulombEwaldDirect:
            v = COULOMBIC_CONVERSION_FACTOR * p[1] * p[2] * math.erfc(p[0] * r) / r
        elif k is Kind.Coulomb:
            v = COULOMBIC_CONVERSION_FACTOR * p[0] * p[1] / r
        elif k is Kind.HardSphere:
            v = math.inf if r < p[0] + p[1] else 0.0
        elif k is Kind.Buckingham:
            v = p[0] * math.exp(-p[1] * r) - p[2] / (r ** 6)
        elif k is Kind.NoInteraction:
            v = 0.0
        elif k is Kind.Monomial:
            v = p[0] / r ** p[1]
        elif k is Kind.Exponential:
            v = p[0] * math.exp(-p[1] * r)
        else:
            raise UndefinedInteractionError()
        return v - self.shift

    def sortkey(self):
        return (int(self.kind), tuple(self.params))

    def same(self, other) -> bool:
        """interactions.jl:257 — equality ignores shift/tailcorrection."""
        return isinstance(other, InteractionRule) and self.kind == other.kind and tuple(self.params) == tuple(other.params)


def make_rule(kind: Kind, params, shift: bool, cutoff: float, tailcorrection: bool | None = None) -> InteractionRule:
    """interactions.jl:350-365 — `InteractionRule(kind, params, shift::Bool, cutoff, tailcorrection)`."""
    if tailcorrection is None:
        tailcorrection = not shift
    params = tuple(float(x) for x in params)
    if shift:
        base = InteractionRule(kind, params, 0.0, False)
        return InteractionRule(kind, params, base(float(cutoff)), tailcorrection)
    return InteractionRule(kind, params, 0.0, tailcorrection)


def NoInteraction() -> InteractionRule:
    return InteractionRule(Kind.NoInteraction, ())


def Undefined() -> InteractionRule:
    return InteractionRule(Kind.UndefinedInteraction, ())


def HardSphere(r1, r2=None) -> InteractionRule:
    return InteractionRule(Kind.HardSphere, (float(r1), float(r1 if r2 is None else r2)))


def LennardJones(eps, sigma) -> InteractionRule:
    return InteractionRule(Kind.LennardJones, (float(eps), float(sigma)))


def Buckingham(A, B, C) -> InteractionRule:
    return InteractionRule(Kind.Buckingham, (float(A), float(B), float(C)))


def CoulombEwaldDirect(alpha, q1, q2=None) -> InteractionRule:
    """interactions.jl:285-303 — collapses to NoInteraction if either charge is zero."""
    q2 = q1 if q2 is None else q2
    if float(q1) == 0.0 or float(q2) == 0.0:
        return NoInteraction()
    return InteractionRule(Kind.CoulombEwaldDirect, (float(alpha), float(q1), float(q2)))


def CoulombRule(q1, q2=None) -> InteractionRule:
    q2 = q1 if q2 is None else q2
    if float(q1) == 0.0 or float(q2) == 0.0:
        return NoInteraction()
    return InteractionRule(Kind.Coulomb, (float(q1), float(q2)))


@dataclass(frozen=True)
class InteractionRuleSum:
    """interactions.jl:558-582 — rules sorted by kind, kinds unique."""
    rules: tuple

    @staticmethod
    def make(rules) -> "InteractionRuleSum":
        rules = list(rules)
        if len(rules) == 0:
            raise ValueError("`InteractionRuleSum([])` is ill-defined.")
        if len(rules) == 1:
            raise ValueError("Defining `InteractionRuleSum([rule])` is forbidden. Directly use `rule` instead.")
        srules = sorted(rules, key=InteractionRule.sortkey)
        if srules[-1].kind is Kind.NoInteraction:
            raise ValueError("Summing any rule with a `NoInteraction` is forbidden.")
        if srules[-1].kind is Kind.UndefinedInteraction:
            raise ValueError("Attempting to sum a rule with an `UndefinedInteraction`.")
        kinds = [r.kind for r in srules]
        if len(set(kinds)) != len(kinds):
            raise ValueError("Repeated rule kind in InteractionRuleSum")
        return InteractionRuleSum(tuple(srules))

    def __call__(self, r):
        ret = 0.0
        for x in self.rules:
            ret += x(r)
        return ret

    def same(self, other) -> bool:
        return isinstance(other, InteractionRuleSum) and len(self.rules) == len(other.rules) and all(
            a.same(b) for a, b in zip(self.rules, other.rules))


Rule = Union[InteractionRule, InteractionRuleSum]


def _rules_of(rule: Rule):
    return (rule,) if isinstance(rule, InteractionRule) else rule.rules


def map_rule(f, rule: Rule) -> Rule:
    """forcefields.jl:176-182."""
    if isinstance(rule, InteractionRule):
        return f(rule)
    return InteractionRuleSum.make([f(r) for r in rule.rules])


def _mix_rules(A: InteractionRule, B: InteractionRule, mixing: Mixing) -> InteractionRule:
    """forcefields.jl:51-98."""
    if B.sortkey() < A.sortkey():
        A, B = B, A
    if B.kind in (Kind.NoInteraction, Kind.UndefinedInteraction):
        return InteractionRule(B.kind, ())
    if B.kind is Kind.HardSphere:
        return HardSphere(A.params[0], B.params[0])
    if A.kind is Kind.HardSphere:
        return HardSphere(A.params[0], 0.0)
    if B.kind in (Kind.CoulombEwaldDirect, Kind.Coulomb):
        if A.kind is not B.kind:
            raise ValueError("Cannot mix cutoff and no-cutoff coulomb interactions")
        if B.kind is Kind.CoulombEwaldDirect:
            if A.params[0] != B.params[0]:
                raise ValueError("Cannot use two Ewald summations with different cutoffs")
            return CoulombEwaldDirect(A.params[0], A.params[1], B.params[1])
        return CoulombRule(A.params[0], B.params[0])
    if A.kind in (Kind.CoulombEwaldDirect, Kind.Coulomb):
        return NoInteraction()
    if B.kind is Kind.LennardJones:
        eA, sA = A.params
        eB, sB = B.params
        if mixing is Mixing.LorentzBerthelot:
            return LennardJones(math.sqrt(eA * eB), (sA + sB) / 2)
        if mixing is Mixing.WaldmanHagler:
            sA3, sB3 = sA ** 3, sB ** 3
            sAB6 = (sA3 * sA3 + sB3 * sB3) / 2
            return LennardJones(math.sqrt(eA * eB) * sA3 * sB3 / sAB6, sAB6 ** (1 / 6))
        if mixing is Mixing.Geometric:
            return LennardJones(math.sqrt(eA * eB), math.sqrt(sA * sB))
        raise AssertionError("no mixing rule")
    if A.kind is not B.kind:
        return Undefined()
    if B.kind is Kind.Buckingham:
        return Buckingham(math.sqrt(A.params[0] * B.params[0]), math.sqrt(A.params[1] * B.params[1]),
                          math.sqrt(A.params[2] * B.params[2]))
    return Undefined()


def reduce_rule_sum(rules: list) -> Rule:
    """forcefields.jl:100-127."""
    rules = sorted(rules, key=InteractionRule.sortkey)
    i = len(rules)
    while i > 0 and rules[i - 1].kind is Kind.NoInteraction:
        i -= 1
    if i == 0:
        return NoInteraction()
    if i == 1:
        return rules[0]
    if rules[i - 1].kind is Kind.UndefinedInteraction:
        return Undefined()
    rules = rules[:i]
    k = 0
    while k < i and rules[k].kind is Kind.HardSphere:
        k += 1
    if k > 0:
        r1, r2 = rules[0].params
        for r in rules[1:k]:
            r1 = max(r1, r.params[0])
            r2 = max(r2, r.params[1])
        newhs = HardSphere(r1, r2)
        if k >= i:
            return newhs
        rules = [newhs] + rules[k:]
    return InteractionRuleSum.make(rules)


def mix_rules(rulei: Rule, rulej: Rule, mixing: Mixing) -> Rule:
    """forcefields.jl:129-174."""
    if isinstance(rulei, InteractionRule) and isinstance(rulej, InteractionRule):
        return _mix_rules(rulei, rulej, mixing)
    if isinstance(rulei, InteractionRule) or isinstance(rulej, InteractionRule):
        rule, rsum = (rulei, rulej) if isinstance(rulei, InteractionRule) else (rulej, rulei)
        if rule.kind is Kind.NoInteraction:
            return rule
        if rule.kind is Kind.UndefinedInteraction:
            return rule
        return reduce_rule_sum([_mix_rules(rule, r, mixing) for r in rsum.rules])
    # sum x sum: the reference's merge loop (forcefields.jl:139-173) never iterates
    # (`while i <= n && m <= j` is false for non-trivial sums), so it reduces an empty list.
    return reduce_rule_sum([])


def sum_rules(r1: Rule, r2: Rule) -> Rule:
    """interactions.jl:612-644."""
    def sum_one(l, r):
        if r.kind is Kind.NoInteraction:
            return list(l)
        if r.kind is Kind.UndefinedInteraction:
            return [r]
        return list(l) + [r]

    if isinstance(r1, InteractionRuleSum):
        if isinstance(r2, InteractionRuleSum):
            return InteractionRuleSum.make(list(r1.rules) + list(r2.rules))
        return InteractionRuleSum.make(sum_one(r1.rules, r2))
    if isinstance(r2, InteractionRuleSum):
        return InteractionRuleSum.make(sum_one(r2.rules, r1))
    if r1.kind is Kind.NoInteraction:
        return r2
    if r1.kind is Kind.UndefinedInteraction:
        return r1
    l = sum_one([r1], r2)
    return l[0] if len(l) == 1 else InteractionRuleSum.make(l)


def get_atom_name(atom) -> str:
    """utils.jl:521-538 — strip a trailing `_123` / `123` numbering."""
    name = str(atom)
    if all(ch.isalpha() for ch in name):
        return name
    s = name.split("_")
    if len(s) > 1:
        if s[-1] != "" and all(ch.isdigit() for ch in s[-1]):
            return "_".join(s[:-1])
    else:
        i = len(name)
        while i > 0 and name[i - 1].isdigit():
            i -= 1
        if i < len(name):
            return name[:i]
    return name


@dataclass
class ForceField:
    """forcefields.jl:5-11.  ``sdict`` maps species name -> 0-based index here."""
    interactions: list            # n x n nested lists of Rule
    sdict: dict
    symbols: list
    cutoff: float
    name: str = "(unnamed)"

    def __getitem__(self, ij) -> Rule:
        i, j = ij
        if isinstance(i, str):
            i, j = self.sdict[i], self.sdict[j]
        return self.interactions[i][j]

    @staticmethod
    def build(input_pairs, mixing: Mixing = Mixing.ErrorOnMix, cutoff: float = 12.0, shift: bool = True,
              tailcorrection: bool | None = None, sdict: dict | None = None, name: str = "(unnamed)") -> "ForceField":
        """forcefields.jl:204-265."""
        if tailcorrection is None:
            tailcorrection = not shift
        if sdict is None:
            allatoms = sorted({a for (a, b), _ in input_pairs} | {b for (a, b), _ in input_pairs})
            sdict = {s: i for i, s in enumerate(allatoms)}
        n = len(sdict)
        symbols = [None] * n
        for s, i in sdict.items():
            symbols[i] = s
        inter = [[None] * n for _ in range(n)]
        done = [[False] * n for _ in range(n)]
        for (a, b), rule in input_pairs:
            i, j = sdict[a], sdict[b]
            if done[i][j] and not inter[i][j].same(rule):
                raise ValueError(f"Interaction between species {a} and {b} defined at multiple occurences")
            inter[i][j] = inter[j][i] = rule
            done[i][j] = done[j][i] = True
        for i in range(n):
            for j in range(i + 1, n):
                if done[i][j]:
                    continue
                if mixing is Mixing.IgnoreInteraction:
                    inter[i][j] = inter[j][i] = NoInteraction()
                elif mixing is Mixing.ErrorOnMix or not done[i][i] or not done[j][j]:
                    inter[i][j] = inter[j][i] = Undefined()
                else:
                    inter[i][j] = inter[j][i] = mix_rules(inter[i][i], inter[j][j], mixing)
        # only i<j pairs are re-wrapped with the global shift/tail flags (forcefields.jl:258-262)
        for i in range(n):
            for j in range(i + 1, n):
                inter[i][j] = inter[j][i] = map_rule(
                    lambda r: make_rule(r.kind, r.params, shift, cutoff, tailcorrection), inter[i][j])
        for i in range(n):
            if inter[i][i] is None:
                inter[i][i] = Undefined()
        return ForceField(inter, dict(sdict), symbols, float(cutoff), name)


def needsvdwgrid(ff: ForceField, atom) -> bool:
    """forcefields.jl:306-316."""
    i = ff.sdict[get_atom_name(atom)]
    for row in ff.interactions:
        for r in _rules_of(row[i]):
            if r.kind not in (Kind.NoInteraction, Kind.CoulombEwaldDirect):
                return True
    return False


# ---------------------------------------------------------------------------------------------
# Flattening for the device: one ceg_pair_t per (probe, framework type) — include/ceg_cuda.h
# ---------------------------------------------------------------------------------------------
PAIR_HS = 1
PAIR_LJ = 2
PAIR_BUCK = 4

pair_dtype = np.dtype([
    ("flags", np.int32), ("_pad", np.int32),
    ("hs_r", np.float64),        # R1 + R2 (Å); the kernel squares it
    ("lj_eps", np.float64), ("lj_sigma", np.float64), ("lj_shift", np.float64),
    ("buck_A", np.float64), ("buck_B", np.float64), ("buck_C", np.float64), ("buck_shift", np.float64),
], align=True)
assert pair_dtype.itemsize == 72


def flatten_pair(rule: Rule) -> np.void:
    """Flatten one entry of ``ForceField.interactions`` as `derivativesGrid` would see it
    (interactions.jl:432-472, 599-610): HardSphere, LJ and Buckingham terms are kept,
    NoInteraction / CoulombEwaldDirect contribute zero, anything else is an error."""
    out = np.zeros((), dtype=pair_dtype)
    for r in _rules_of(rule):
        k = r.kind
        if k in (Kind.NoInteraction, Kind.CoulombEwaldDirect):
            continue
        if k is Kind.HardSphere:
            out["flags"] |= PAIR_HS
            out["hs_r"] = r.params[0] + r.params[1]
            if r.shift != 0.0:
                raise ValueError("shifted HardSphere is not representable")
        elif k is Kind.LennardJones:
            out["flags"] |= PAIR_LJ
            out["lj_eps"], out["lj_sigma"] = r.params
            out["lj_shift"] = r.shift
        elif k is Kind.Buckingham:
            out["flags"] |= PAIR_BUCK
            out["buck_A"], out["buck_B"], out["buck_C"] = r.params
            out["buck_shift"] = r.shift
        elif k is Kind.Coulomb:
            raise ValueError("Coulomb interactions should not be taken into account in VdW grids.")
        elif k is Kind.UndefinedInteraction:
            raise UndefinedInteractionError()
        else:
            raise ValueError(f"VdW grid not implemented for {k.name}")
    return out


def pair_table(ff: ForceField, probes, n_types: int | None = None) -> np.ndarray:
    """(n_probes, n_types) table of ceg_pair_t: row p = interactions of probe species p
    with every force-field species (probes.jl:84 indexes rule[atomkind, probe])."""
    n = len(ff.sdict) if n_types is None else n_types
    tab = np.zeros((len(probes), n), dtype=pair_dtype)
    for p, probe in enumerate(probes):
        ip = ff.sdict[get_atom_name(probe)] if isinstance(probe, str) else int(probe)
        for t in range(n):
            try:
                tab[p, t] = flatten_pair(ff.interactions[t][ip])
            except UndefinedInteractionError:
                tab[p, t]["flags"] = -1   # raises only if a framework atom of that type exists
    return tab
