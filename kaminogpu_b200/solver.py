"""Python mirror of the reference's solver surface over the kamino_b200 C ABI.

``KaminoSolver`` / ``KaminoQuantity`` / ``KaminoParticles`` keep the reference's names,
argument meaning and call order (include/KaminoSolver.cuh:99-111,
include/KaminoQuantity.cuh:38-69, include/KaminoParticles.cuh:6-24 under
/root/reference/KaminoGPU/), so the parity tests read like a driver of the reference.
All device work goes through ``libkamino_b200.so``; numpy arrays are host mirrors only.
The C++ classes in ``kaminogpu_b200/host/`` are the compiled drop-in; this module exists
for tests and bench.py.
"""
import ctypes

import numpy as np

from . import capi

M_PI = 3.14159265358979323846
M_2PI = 6.28318530717958647692


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class KaminoQuantity:
    """Host mirror + device view of one field (kernel/KaminoQuantity.cu)."""

    def __init__(self, solver, field, name, nPhi, nTheta, phiOffset, thetaOffset, sim=0):
        self._solver, self._field, self._sim = solver, field, sim
        self.attrName = name
        self.nPhi, self.nTheta = nPhi, nTheta
        self.phiOffset, self.thetaOffset = phiOffset, thetaOffset
        self.cpuBuffer = np.zeros((nTheta, nPhi), dtype=np.float32)   # [theta][phi], KaminoQuantity.cu:70-73

    def getName(self):
        return self.attrName

    def getNPhi(self):
        return self.nPhi

    def getNTheta(self):
        return self.nTheta

    def getPhiOffset(self):
        return self.phiOffset

    def getThetaOffset(self):
        return self.thetaOffset

    def getCPUValueAt(self, phi, theta):
        return float(self.cpuBuffer[theta, phi])

    def setCPUValueAt(self, phi, theta, val):
        self.cpuBuffer[theta, phi] = val

    def copyToGPU(self):
        buf = np.ascontiguousarray(self.cpuBuffer, dtype=np.float32)
        capi.check(self._solver._lib.kamino_upload_field(self._solver._ctx, self._field, self._sim, _ptr(buf)),
                   self._solver._ctx)

    def copyBackToCPU(self):
        capi.check(self._solver._lib.kamino_download_field(self._solver._ctx, self._field, self._sim,
                                                           _ptr(self.cpuBuffer)), self._solver._ctx)
        return self.cpuBuffer

    def _device(self, which):
        p, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        capi.check(self._solver._lib.kamino_field_device_ptr(self._solver._ctx, self._field, self._sim, which,
                                                             ctypes.byref(p), ctypes.byref(pitch)), self._solver._ctx)
        return p.value, pitch.value

    def getGPUThisStep(self):
        return self._device(0)[0]

    def getGPUNextStep(self):
        return self._device(1)[0]

    def getThisStepPitchInElements(self):
        return self._device(0)[1]

    def getNextStepPitchInElements(self):
        return self._device(1)[1]


class KaminoParticles:
    """Tracer particles (kernel/KaminoParticles.cu): jittered lattice seeded with libc rand()."""

    def __init__(self, solver, particleDensity, gridLen, nTheta, sim=0, coords=None):
        self._solver, self._sim = solver, sim
        self.nTheta, self.nPhi = nTheta, 2 * nTheta
        lib = solver._lib
        if coords is None:
            self.numOfParticles = int(lib.kamino_particle_count(nTheta, ctypes.c_float(particleDensity)))
            self.coordCPUBuffer = np.zeros(2 * self.numOfParticles, dtype=np.float32)
            if self.numOfParticles:
                capi.check(lib.kamino_seed_particles_host(nTheta, ctypes.c_float(particleDensity),
                                                          _ptr(self.coordCPUBuffer)))
        else:
            self.coordCPUBuffer = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1).copy()
            self.numOfParticles = self.coordCPUBuffer.size // 2
        self.colorBGR = np.zeros(3 * self.numOfParticles, dtype=np.float32)   # no image: black, :73-77

    def copy2GPU(self):
        capi.check(self._solver._lib.kamino_upload_particles(self._solver._ctx, self._sim,
                                                             _ptr(self.coordCPUBuffer)), self._solver._ctx)

    def copyBack2CPU(self):
        capi.check(self._solver._lib.kamino_download_particles(self._solver._ctx, self._sim,
                                                               _ptr(self.coordCPUBuffer)), self._solver._ctx)
        return self.coordCPUBuffer


class KaminoSolver:
    """KaminoSolver(nPhi, nTheta, radius, frameDuration, A, B, C, D, E)
    (kernel/KaminoSolver.cu:12-67). As in Kamino::run (kernel/KaminoCore.cu:862) the
    fourth argument receives dt, which is also the time step every kernel uses.
    ``batch`` > 1 creates an ensemble of identical-shape simulations in one context."""

    def __init__(self, nPhi, nTheta, radius, frameDuration, A=0.0, B=1, C=1, D=1, E=1,
                 device=0, batch=1, initVelocity=True):
        if nPhi != 2 * nTheta:
            raise ValueError("nPhi must be 2 * nTheta (kernel/KaminoCore.cu:849)")
        self._lib = capi.load()
        self.nPhi, self.nTheta, self.radius = nPhi, nTheta, float(radius)
        self.gridLen = np.float32(M_2PI / nPhi)
        self.frameDuration = float(frameDuration)
        self.batch = batch
        self.timeStep = 0.0
        self.timeElapsed = 0.0
        ctx = ctypes.c_void_p()
        capi.check(self._lib.kamino_create(ctypes.byref(ctx), device, nTheta, ctypes.c_float(radius),
                                           ctypes.c_float(frameDuration), batch, 0))
        self._ctx = ctx
        self.velPhi = KaminoQuantity(self, capi.VEL_PHI, "velPhi", nPhi, nTheta, -0.5, 0.5)
        self.velTheta = KaminoQuantity(self, capi.VEL_THETA, "velTheta", nPhi, nTheta - 1, 0.0, 1.0)
        self.pressure = KaminoQuantity(self, capi.PRESSURE, "p", nPhi, nTheta, 0.0, 0.5)
        self.density = KaminoQuantity(self, capi.DENSITY, "density", nPhi, nTheta, 0.0, 0.5)
        self.particles = None
        if initVelocity:
            self.initialize_velocity()

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.kamino_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- initialisation ------------------------------------------------------------------
    def quantity(self, field, sim):
        """View of `field` of simulation `sim` of an ensemble."""
        proto = {capi.VEL_PHI: self.velPhi, capi.VEL_THETA: self.velTheta,
                 capi.PRESSURE: self.pressure, capi.DENSITY: self.density}[field]
        return KaminoQuantity(self, field, proto.attrName, proto.nPhi, proto.nTheta,
                              proto.phiOffset, proto.thetaOffset, sim=sim)

    def initialize_velocity(self):
        """kernel/KaminoInitializer.cu:3-85 (every simulation of an ensemble gets the same field)."""
        capi.check(self._lib.kamino_init_velocity_host(self.nTheta, ctypes.c_float(self.radius),
                                                       _ptr(self.velPhi.cpuBuffer), _ptr(self.velTheta.cpuBuffer)))
        for sim in range(self.batch):
            for q in (self.velPhi, self.velTheta):
                capi.check(self._lib.kamino_upload_field(self._ctx, q._field, sim, _ptr(q.cpuBuffer)), self._ctx)

    def initialize_velocity_on_device(self):
        """The same FBM field evaluated by a CUDA kernel (kamino_init_velocity_device): start-up of large grids."""
        capi.check(self._lib.kamino_init_velocity_device(self._ctx), self._ctx)

    def seed_particles_on_device(self, particleDensity, seed):
        """The reference's particle lattice with counter-based jitter (kamino_seed_particles_device)."""
        n = int(self._lib.kamino_particle_count(self.nTheta, ctypes.c_float(particleDensity)))
        self.particles = KaminoParticles(self, 0.0, self.gridLen, self.nTheta, coords=np.zeros(2 * n, np.float32))
        capi.check(self._lib.kamino_alloc_particles(self._ctx, n), self._ctx)
        capi.check(self._lib.kamino_seed_particles_device(self._ctx, ctypes.c_float(particleDensity),
                                                          ctypes.c_ulonglong(seed)), self._ctx)

    def initDensityfromPic(self, path):
        """kernel/KaminoSolver.cu:243-277: a no-op for "" (image input is out of scope)."""
        if path:
            raise NotImplementedError("density images are out of scope (SURVEY.md section 8f)")

    def initParticlesfromPic(self, path, parPerGrid, coords=None):
        """kernel/KaminoSolver.cu:279-282; parPerGrid is truncated to an integer like the
        reference's size_t parameter. `coords` supplies explicit positions instead."""
        if path:
            raise NotImplementedError("particle colour images are out of scope (SURVEY.md section 8f)")
        self.particles = KaminoParticles(self, float(int(parPerGrid)), self.gridLen, self.nTheta, coords=coords)
        capi.check(self._lib.kamino_alloc_particles(self._ctx, self.particles.numOfParticles), self._ctx)
        for sim in range(self.batch):
            capi.check(self._lib.kamino_upload_particles(self._ctx, sim, _ptr(self.particles.coordCPUBuffer)),
                       self._ctx)

    # -- the hot path ----------------------------------------------------------------------
    def advection(self):
        capi.check(self._lib.kamino_advect(self._ctx), self._ctx)

    def geometric(self):
        capi.check(self._lib.kamino_geometric(self._ctx), self._ctx)

    def projection(self):
        capi.check(self._lib.kamino_project(self._ctx), self._ctx)

    def projection_cr_order(self):
        """Parity instrumentation: the projection with the reference's cyclic-reduction theta solve."""
        capi.check(self._lib.kamino_debug_project_cr(self._ctx), self._ctx)

    def stepForward(self, timeStep=None, nSteps=1):
        """kernel/KaminoSolver.cu:197-221. The argument is recorded and otherwise ignored,
        exactly as the reference's kernels ignore it."""
        self.timeStep = self.frameDuration if timeStep is None else timeStep
        capi.check(self._lib.kamino_step(self._ctx, nSteps), self._ctx)
        self.timeElapsed += self.timeStep * nSteps

    def sync(self):
        capi.check(self._lib.kamino_sync(self._ctx), self._ctx)

    def phase_times(self, reset=False):
        a, g, p = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        capi.check(self._lib.kamino_phase_times(self._ctx, ctypes.byref(a), ctypes.byref(g), ctypes.byref(p),
                                                int(reset)), self._ctx)
        return a.value, g.value, p.value

    def set_stream(self, cuda_stream_ptr):
        capi.check(self._lib.kamino_set_stream(self._ctx, ctypes.c_void_p(cuda_stream_ptr)), self._ctx)

    def locate(self, kind, phiRaw, thetaRaw):
        """Device evaluation of the samplers' index / predicate logic (parity instrumentation)."""
        phi = np.ascontiguousarray(phiRaw, dtype=np.float32)
        theta = np.ascontiguousarray(thetaRaw, dtype=np.float32)
        n = phi.size
        out = {k: np.zeros(n, dtype=t) for k, t in (("phiIndex", np.int32), ("thetaIndex", np.int32),
                                                     ("alphaPhi", np.float32), ("alphaTheta", np.float32),
                                                     ("phi", np.float32), ("theta", np.float32),
                                                     ("flags", np.int32))}
        capi.check(self._lib.kamino_debug_locate(self._ctx, kind, n, _ptr(phi), _ptr(theta),
                                                 _ptr(out["phiIndex"]), _ptr(out["thetaIndex"]),
                                                 _ptr(out["alphaPhi"]), _ptr(out["alphaTheta"]),
                                                 _ptr(out["phi"]), _ptr(out["theta"]), _ptr(out["flags"])), self._ctx)
        return out


def load_config(path):
    """configKamino.txt grammar (kernel/main.cu:17-45): 16 whitespace-separated tokens."""
    tok = open(path).read().split()
    if len(tok) < 16:
        raise ValueError("configKamino.txt needs 16 tokens, found %d" % len(tok))
    names = ["radius", "nTheta", "particleDensity", "dt", "DT", "frames", "A", "B", "C", "D", "E",
             "gridPath", "particlePath", "densityImage", "solidImage", "colorImage"]
    types = [float, int, float, float, float, int, float, int, int, int, int, str, str, str, str, str]
    cfg = {n: t(v) for n, t, v in zip(names, types, tok)}
    for key in ("densityImage", "colorImage"):       # the solid image's "null" is left as is (:37-40)
        if cfg[key] == "null":
            cfg[key] = ""
    return cfg


def steps_per_frame(dt, DT, frames):
    """Number of stepForward calls Kamino::run makes for each frame (kernel/KaminoCore.cu:886-895):
    float arithmetic, the `while (T < i*DT)` iterations plus the always-taken remainder step."""
    dt, DT = np.float32(dt), np.float32(DT)
    T = np.float32(0.0)
    out = []
    for i in range(1, frames + 1):
        n = 0
        while T < np.float32(i) * DT:
            T = np.float32(T + dt)
            n += 1
        n += 1
        T = np.float32(i) * DT
        out.append(n)
    return out
