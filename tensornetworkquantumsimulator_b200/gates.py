"""Gate tuples -> dense ComplexF64 gate arrays (host side).

Mirrors `/root/reference/src/Apply/gate_definitions.jl`: `toitensor` :4-6,34-57, `param_rescaling`
:28-31 (Rxx/Ryy/Rzz tuples carry qiskit's angle, the ITensors op takes half of it), the custom ops
`xx_plus_yy` :98-105, `Rxxyy` :112-115, `Rxxyyzz` :122-125.  The ITensors "Qubit"/"S=1/2" operator
matrices are not in /root/reference; they are restated here.

Array conventions (also the C-ABI's): a 1-site gate is o[s', s]; a 2-site gate is
o[s1', s2', s1, s2] (first vertex of the tuple = index 1; Kronecker order, first site slow, when
viewed as a d^2 x d^2 matrix).  T'[s'] = sum_s o[s', s] T[s]  (ITensors.apply).
"""
from __future__ import annotations

import numpy as np

I2 = np.eye(2, dtype=np.complex128)
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
PAULI = {"I": I2, "Id": I2, "X": X, "Y": Y, "Z": Z}

_SQ = 1.0 / np.sqrt(2.0)
STATES = {
    "↑": [1, 0], "Up": [1, 0], "0": [1, 0], "Z+": [1, 0], "Zp": [1, 0],
    "↓": [0, 1], "Dn": [0, 1], "1": [0, 1], "Z-": [0, 1], "Zm": [0, 1],
    "X+": [_SQ, _SQ], "+": [_SQ, _SQ], "X-": [_SQ, -_SQ], "-": [_SQ, -_SQ],
    "Y+": [_SQ, 1j * _SQ], "Y-": [_SQ, -1j * _SQ],
}


def state_vector(name) -> np.ndarray:
    if isinstance(name, str):
        return np.asarray(STATES[name], dtype=np.complex128)
    return np.asarray(name, dtype=np.complex128)


def _expm_herm(h: np.ndarray, theta: float) -> np.ndarray:
    w, v = np.linalg.eigh(h)
    return (v * np.exp(-1j * theta * w)) @ v.conj().T


def param_rescaling(name: str, param):
    """/root/reference/src/Apply/gate_definitions.jl:28-31."""
    if name in ("Rxx", "Ryy", "Rzz"):
        return param / 2
    return param


def op_matrix(name: str, param=None) -> np.ndarray:
    """Operator matrix in ITensors' qubit conventions; `param` is the op's own θ/ϕ (already rescaled)."""
    c = s = None
    if param is not None and not isinstance(param, (tuple, list)):
        c, s = np.cos(param / 2), np.sin(param / 2)
    if name in PAULI:
        return PAULI[name].copy()
    if name == "H":
        return np.array([[1, 1], [1, -1]], dtype=np.complex128) * _SQ
    if name in ("S", "Phase") and param is None:
        return np.diag([1, 1j]).astype(np.complex128)
    if name == "T":
        return np.diag([1, np.exp(0.25j * np.pi)]).astype(np.complex128)
    if name in ("P", "Phase"):
        return np.diag([1, np.exp(1j * param)]).astype(np.complex128)
    if name == "Rx":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "Ry":
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == "Rz":
        return np.diag([np.exp(-0.5j * param), np.exp(0.5j * param)]).astype(np.complex128)
    if name in ("CNOT", "CX"):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    if name == "CY":
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]], dtype=np.complex128)
    if name == "CZ":
        return np.diag([1, 1, 1, -1]).astype(np.complex128)
    if name == "SWAP":
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    if name == "iSWAP":
        return np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    if name == "CPHASE":
        return np.diag([1, 1, 1, np.exp(1j * param)]).astype(np.complex128)
    if name == "Rxx":   # exp(-i ϕ XX)
        return np.cos(param) * np.eye(4) - 1j * np.sin(param) * np.kron(X, X)
    if name == "Ryy":
        return np.cos(param) * np.eye(4) - 1j * np.sin(param) * np.kron(Y, Y)
    if name == "Rzz":
        return np.diag(np.exp(-1j * param * np.array([1, -1, -1, 1]))).astype(np.complex128)
    if name in ("CRx", "CRy", "CRz"):
        u = np.eye(4, dtype=np.complex128)
        u[2:, 2:] = op_matrix(name[1:], param)
        return u
    if name == "Rxxyy":       # gate_definitions.jl:112-115
        return _expm_herm(0.5 * (np.kron(X, X) + np.kron(Y, Y)), param)
    if name == "Rxxyyzz":     # gate_definitions.jl:122-125
        return _expm_herm(0.5 * (np.kron(X, X) + np.kron(Y, Y) + np.kron(Z, Z)), param)
    if name == "xx_plus_yy":  # gate_definitions.jl:98-105, param = (θ, β)
        th, be = param
        return np.array([
            [1, 0, 0, 0],
            [0, np.cos(th / 2), -1j * np.sin(th / 2) * np.exp(-1j * be), 0],
            [0, -1j * np.sin(th / 2) * np.exp(1j * be), np.cos(th / 2), 0],
            [0, 0, 0, 1]], dtype=np.complex128)
    raise ValueError(f"Wrong gate format: unknown gate {name!r}")


# ---- Heisenberg picture: "Pauli" sites (d = 4, basis I, X, Y, Z) and Pauli transfer matrices ---------------------------------
PAULI_STATES = {"I": [1, 0, 0, 0], "X": [0, 1, 0, 0], "Y": [0, 0, 1, 0], "Z": [0, 0, 0, 1]}   # tensornetworkstate_constructors.jl:1
_PAULI_BASIS = (I2, X, Y, Z)


def pauli_rotation_matrix(generator: str, theta: float) -> np.ndarray:
    """exp(-i theta/2 P) for the Pauli string P = generator (first letter = first vertex = slow Kronecker factor);
    `paulirotationmatrix`, gate_definitions.jl:60-68 (PauliPropagation.tomatrix of a PauliRotation: that package is not in
    /root/reference; its published convention R_P(theta) = exp(-i theta/2 P) is restated here)."""
    p = np.array([[1.0 + 0j]])
    for ch in generator.upper():
        p = np.kron(p, PAULI[ch])
    return np.cos(theta / 2) * np.eye(p.shape[0]) - 1j * np.sin(theta / 2) * p


def pauli_transfer_matrix(u: np.ndarray, heisenberg: bool = True) -> np.ndarray:
    """Pauli transfer matrix of the unitary u on n qubits as an array M[a_1..a_n, b_1..b_n] (a_i, b_i in I, X, Y, Z):
    heisenberg: U^dag P_b U = sum_a M[a, b] P_a, i.e. M[a, b] = tr(P_a U^dag P_b U) / 2^n (an operator with Pauli coefficients c
    becomes M c); otherwise the Schroedinger map U P_b U^dag.  Real for unitary u.  `PP.calculateptm(U; heisenberg = true)`
    at gate_definitions.jl:77 (PauliPropagation is not in /root/reference: restated from its documented definition)."""
    n = int(round(np.log2(u.shape[0])))
    w = u if heisenberg else u.conj().T
    strings = [np.array([[1.0 + 0j]])]
    for _ in range(n):
        strings = [np.kron(s, q) for s in strings for q in _PAULI_BASIS]          # first factor slow: index (a_1, .., a_n) row-major
    m = np.empty((4 ** n, 4 ** n), dtype=np.complex128)
    for b, pb in enumerate(strings):
        img = w.conj().T @ pb @ w
        for a, pa in enumerate(strings):
            m[a, b] = np.trace(pa @ img) / 2 ** n
    return m.reshape((4,) * (2 * n))


def heisenberg_gate_array(gate: tuple) -> np.ndarray:
    """("R" + Pauli string, vertices, theta) on Pauli sites -> o[s', s] / o[s1', s2', s1, s2] with d = 4
    (`toitensor_heisenberg`, gate_definitions.jl:70-86: the angle is NOT rescaled on this branch, :41-42)."""
    name = gate[0]
    nverts = len(gate[1]) if isinstance(gate[1], list) else 1
    if not (name[:1] == "R" and _is_pauli_string(name[1:]) and len(name) - 1 == nverts and len(gate) >= 3):
        raise ValueError(f"Wrong gate format: Pauli sites take ('R<pauli string>', vertices, theta), got {name!r}")
    # real for a unitary: rounding residue of the imaginary part is dropped
    return pauli_transfer_matrix(pauli_rotation_matrix(name[1:], gate[2]), heisenberg=True).real.astype(np.complex128)


def _is_pauli_string(s: str) -> bool:
    return len(s) > 0 and all(ch in "XYZIxyz" for ch in s)


def gate_array(gate: tuple, site_dims=None) -> np.ndarray:
    """(name, vertices[, param]) -> o[s', s] or o[s1', s2', s1, s2] (`toitensor`, gate_definitions.jl:34-57).
    site_dims = the dimensions of the sites the gate acts on: all 4 ("Pauli" sites) selects the Heisenberg branch (:41-42)."""
    if site_dims is not None and len(site_dims) > 0 and all(int(x) == 4 for x in site_dims):
        return heisenberg_gate_array(gate)
    name = gate[0]
    nverts = len(gate[1]) if isinstance(gate[1], list) else 1
    if _is_pauli_string(name) and len(name) == nverts and name not in ("I",) or (name in PAULI and nverts == 1):
        mats = [PAULI[ch.upper()] for ch in name] if name not in PAULI else [PAULI[name]]
        u = mats[0]
        for m in mats[1:]:
            u = np.kron(u, m)
    elif len(gate) == 2:
        u = op_matrix(name)
    else:
        u = op_matrix(name, param_rescaling(name, gate[2]))
    d = 2
    if u.shape == (d, d):
        if nverts != 1:
            raise ValueError("one-site gate given two vertices")
        return u.astype(np.complex128)
    if nverts != 2:
        raise ValueError("two-site gate needs two vertices")
    return u.astype(np.complex128).reshape(d, d, d, d)
