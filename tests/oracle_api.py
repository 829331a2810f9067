"""ctypes access to the CPU oracle (oracle/kamino_oracle.c) -- test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SRC = os.path.join(ROOT, "oracle", "kamino_oracle.c")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libkamino_oracle.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

VPHI, VTHETA, CENTERED = 0, 1, 2


class Params(ctypes.Structure):
    _fields_ = [("nTheta", ctypes.c_int), ("nPhi", ctypes.c_int), ("radius", ctypes.c_float),
                ("dt", ctypes.c_float), ("gridLen", ctypes.c_float)]


class Location(ctypes.Structure):
    _fields_ = [("phiIndex", ctypes.c_int), ("thetaIndex", ctypes.c_int), ("alphaPhi", ctypes.c_float),
                ("alphaTheta", ctypes.c_float), ("phi", ctypes.c_float), ("theta", ctypes.c_float),
                ("flipped", ctypes.c_int), ("poleBranch", ctypes.c_int)]


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(ORACLE_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               "-o", ORACLE_LIB, ORACLE_SRC, "-lm"])
    return ORACLE_LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
        _lib.ko_particle_count.restype = ctypes.c_long
        _lib.ko_sample.restype = ctypes.c_float
    return _lib


def fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def params(nTheta, radius=5.0, dt=0.005):
    return Params(nTheta, 2 * nTheta, np.float32(radius), np.float32(dt), np.float32(np.pi / nTheta))


def step(p, velPhi, velTheta, density, particles, phase=0):
    """Advance copies of the state by one step (phase 1/2 = stop after advection/geometric)."""
    u, v = velPhi.copy(), velTheta.copy()
    rho = density.copy() if density is not None else None
    pc = particles.copy() if particles is not None else None
    pressure = np.zeros(p.nTheta * p.nPhi, dtype=np.float32)
    n = 0 if pc is None else pc.size // 2
    lib().ko_step(ctypes.byref(p), fptr(u), fptr(v), fptr(rho), fptr(pressure), fptr(pc), ctypes.c_long(n), phase)
    return u, v, rho, pc, pressure


def geometric(p, velPhi, velTheta):
    uo, vo = np.zeros_like(velPhi), np.zeros_like(velTheta)
    su, sv = np.zeros_like(velPhi), np.zeros_like(velPhi)
    lib().ko_geometric(ctypes.byref(p), fptr(velPhi), fptr(velTheta), fptr(uo), fptr(vo), fptr(su), fptr(sv))
    return uo, vo


def projection(p, velPhi, velTheta):
    u, v = velPhi.copy(), velTheta.copy()
    pressure = np.zeros(p.nTheta * p.nPhi, dtype=np.float32)
    lib().ko_projection(ctypes.byref(p), fptr(u), fptr(v), fptr(pressure))
    return u, v, pressure


def locate(p, kind, phi, theta):
    loc = Location()
    lib().ko_locate(ctypes.byref(p), kind, ctypes.c_float(phi), ctypes.c_float(theta), ctypes.byref(loc))
    return loc


def init_velocity(nTheta, radius=5.0):
    u = np.zeros(nTheta * 2 * nTheta, np.float32)
    v = np.zeros((nTheta - 1) * 2 * nTheta, np.float32)
    lib().ko_init_velocity(nTheta, ctypes.c_float(radius), fptr(u), fptr(v))
    return u, v


def synthetic_density(nTheta):
    rho = np.zeros(nTheta * 2 * nTheta, np.float32)
    lib().ko_synthetic_density(nTheta, fptr(rho))
    return rho


def seed_particles(nTheta, density):
    n = lib().ko_particle_count(nTheta, ctypes.c_float(density))
    pc = np.zeros(2 * n, np.float32)
    if n:
        lib().ko_seed_particles(nTheta, ctypes.c_float(density), fptr(pc))
    return pc


def golden(case):
    return np.load(os.path.join(GOLDEN_DIR, "ref_%s.npz" % case))


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
