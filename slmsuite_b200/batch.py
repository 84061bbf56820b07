"""
``HologramBatch``: B independent holograms of identical ``shape`` / ``slm_shape`` optimised in one
launch sequence (``blockIdx.y`` = hologram), and the multi-GPU sharding of a batch.

The reference has no batch: "a batch" there is a Python list of ``Hologram`` objects run one after
the other.  Each hologram's loop is independent, so a batch shards across GPUs by
hologram with no communication inside the loop and ONE all-gather of the final phases at the end
(SURVEY.md 8e).  That collective is NCCL behind the C ABI (``slmgs_allgather_phase``, NCCL loaded with dlopen inside
``libslmgs.so``); the rendezvous is ``slmsuite_b200.comm`` -- no torch anywhere in the package.
"""

import numpy as np

from . import _lib
from .hologram import Hologram, _norm


def shard_bounds(n_items, rank, world):
    """Contiguous block partition: rank r owns [lo, hi) with ceil(n/world) items per rank (last may be short)."""
    per = -(-int(n_items) // int(world))
    lo = min(rank * per, n_items)
    hi = min(lo + per, n_items)
    return lo, hi


class HologramBatch(Hologram):
    """
    ``HologramBatch(targets, amp=None, phase=None, slm_shape=None, **flags)``

    targets : (B, H, W) array (or (H, W) with ``batch=B`` to share one target)
    phase   : (B, h, w) initial phases (or None: independent random phases)
    amp     : None (uniform), (h, w) shared, or (B, h, w)
    All holograms share the method / flags of ``optimize()``; state accessors return arrays with a
    leading batch axis.
    """

    def __init__(self, targets, amp=None, phase=None, slm_shape=None, dtype=np.float32,
                 propagation_kernel=None, device=0, batch=None, **kwargs):
        targets = np.asarray(targets)
        if targets.ndim == 2:
            if batch is None:
                raise ValueError("a shared (H, W) target needs batch=")
            self._B = int(batch)
            self._shared_target = True
        elif targets.ndim == 3:
            self._B = int(targets.shape[0])
            self._shared_target = False
        else:
            raise ValueError(f"Unexpected targets of shape {targets.shape}.")
        if self._B < 1:
            raise ValueError("empty batch")
        self._pending_targets = targets
        ph = None if phase is None else np.asarray(phase, dtype=np.float32)
        if ph is not None and ph.ndim == 2:
            ph = np.broadcast_to(ph, (self._B,) + ph.shape)
        if ph is not None and ph.shape[0] != self._B:
            raise ValueError(f"phase batch {ph.shape[0]} does not match targets batch {self._B}")
        self._pending_phase = ph
        am = None if amp is None else np.asarray(amp, dtype=np.float32)
        first_target = targets if targets.ndim == 2 else targets[0]
        super().__init__(first_target, amp=None if am is None else (am if am.ndim == 2 else am[0]),
                         phase=None if ph is None else ph[0], slm_shape=slm_shape, dtype=dtype,
                         propagation_kernel=propagation_kernel, device=device, **kwargs)
        if am is not None and am.ndim == 3:
            if am.shape[0] != self._B:
                raise ValueError(f"amp batch {am.shape[0]} does not match targets batch {self._B}")
            a = np.array(am, dtype=np.float32)
            for b in range(self._B):
                a[b] *= 1 / _norm(a[b])
            self._amp = a
            self._check(self._lib.slmgs_set_amp_array(self._ctx, _lib.fptr(_lib.f32(a)), 1))

    def _batch_size(self):
        return self._B

    def _bshape(self, shape):
        return (self._B,) + tuple(shape)

    def __len__(self):
        return self._B

    # ---- batched state ------------------------------------------------------------------------
    def _set_target(self, new_target, reset_weights=False):
        if getattr(self, "_pending_targets", None) is not None:
            new_target = self._pending_targets
            self._pending_targets = None
        if new_target is None:
            t = np.zeros(self.shape, dtype=self.dtype)
            self._shared_target = True
        else:
            t = np.array(new_target, dtype=self.dtype)
            if t.shape[-2:] != tuple(self.shape):
                raise ValueError(f"Target shape {t.shape} does not match hologram shape {self.shape}")
            self._shared_target = t.ndim == 2
            np.abs(t, out=t)
            with np.errstate(all="ignore"):
                if t.ndim == 2:
                    t *= 1 / _norm(t)
                else:
                    if t.shape[0] != self._B:
                        raise ValueError(f"targets batch {t.shape[0]} does not match batch {self._B}")
                    for b in range(self._B):
                        t[b] *= 1 / _norm(t[b])
        self._target = t
        self._upload_target()
        if reset_weights:
            self.reset_weights()

    def _upload_target(self):
        self._mraf_cache = None
        self._check(self._lib.slmgs_set_target(self._ctx, _lib.fptr(_lib.f32(self._target)),
                                               1 if self._shared_target else 0))

    def reset_phase(self, custom_phase=None, random_phase=None, quadratic_phase=None):
        if getattr(self, "_pending_phase", None) is not None:
            custom_phase = self._pending_phase
            self._pending_phase = None
        elif custom_phase is not None and not self._phase_set:
            custom_phase = None  # constructor passed the first hologram's slice of a consumed batch
        if custom_phase is not None:
            p = np.array(custom_phase, dtype=self.dtype)
            if p.ndim == 2:
                p = np.broadcast_to(p, (self._B,) + p.shape)
            if p.shape != (self._B,) + tuple(self.slm_shape):
                raise ValueError(f"Reset phase of shape {p.shape} is not of slm_shape {self.slm_shape}")
        else:
            if quadratic_phase is None:
                quadratic_phase = self.flags.get("quadratic_phase", False)
            if quadratic_phase:
                raise NotImplementedError("quadratic_phase preconditioning is outside the GS/WGS hot path; pass phase=")
            if random_phase is None:
                random_phase = self.flags.get("random_phase", 1)
            rng = np.random.default_rng()
            p = (random_phase * rng.uniform(-np.pi, np.pi, (self._B,) + tuple(self.slm_shape))).astype(self.dtype)
        self._check(self._lib.slmgs_set_phase(self._ctx, _lib.fptr(_lib.f32(p))))
        self._phase_set = True

    def set_weights(self, new_weights):
        w = np.asarray(new_weights)
        if w.shape != (self._B,) + tuple(self.shape):
            raise ValueError(f"New weights {w.shape} do not match target shape {(self._B,) + tuple(self.shape)}")
        self._check(self._lib.slmgs_set_weights(self._ctx, _lib.fptr(_lib.f32(w))))

    @property
    def nearfield(self):
        raise NotImplementedError("nearfield is rebuilt per hologram: use Hologram for that accessor")

    def _iteration_params(self, mraf, stepped):
        if self.flags.get("fix_phase_efficiency", None) is not None:
            raise NotImplementedError("fix_phase_efficiency is per hologram; a batch shares one flag state")
        return super()._iteration_params(mraf, stepped)

    def _calculate_stats_computational(self, stats, stat_groups=[]):
        if "computational" in stat_groups:
            per = self._stats_pixel()
            stats["computational"] = {k: np.array([d[k] for d in per]) for k in per[0]}

    # ---- multi-GPU ----------------------------------------------------------------------------
    def gather_phases(self, n_total=None, comm=None):
        """
        All-gather of the final near-field phases over the job's ranks (one collective per job): NCCL, loaded inside
        ``libslmgs.so`` (``slmgs_allgather_phase``); the rendezvous is ``slmsuite_b200.comm`` (no torch).  Returns a
        host array (n_total, h, w) on every rank.  A single process just returns ``self.phase``.
        ``comm``: a ``slmsuite_b200.comm.Comm`` (default: the one built from the launcher's environment), or any object
        with the same ``allgather_phase(holo, n_total)`` method.
        """
        from . import comm as _comm

        comm = comm if comm is not None else _comm.default()
        if comm.world == 1:
            return self.phase
        return comm.allgather_phase(self, n_total=n_total)


def optimize_sharded(targets, phases, method="GS", maxiter=20, slm_shape=None, amp=None, device=None, comm=None,
                     **kwargs):
    """
    Optimise a batch of independent holograms sharded over the ranks of a job (one process per GPU): rank r builds a
    ``HologramBatch`` of its contiguous block, runs the loop with no communication, and all ranks receive all final
    phases from ONE all-gather.  Returns (phases (B, h, w), local HologramBatch or None if this rank owns nothing).
    """
    from . import comm as _comm

    comm = comm if comm is not None else _comm.default()
    targets = np.asarray(targets)
    phases = np.asarray(phases)
    n = phases.shape[0]
    rank, world = comm.rank, comm.world
    lo, hi = shard_bounds(n, rank, world)
    if device is None:
        device = getattr(comm, "device", rank)
    holo = None
    if hi > lo:
        t = targets if targets.ndim == 2 else targets[lo:hi]
        a = amp if (amp is None or np.ndim(amp) == 2) else np.asarray(amp)[lo:hi]
        holo = HologramBatch(t, amp=a, phase=phases[lo:hi], slm_shape=slm_shape, device=device, batch=hi - lo)
        holo.optimize(method, maxiter=maxiter, verbose=False, **kwargs)
        if world == 1:
            return holo.phase, holo
    # a rank without work still takes part in the collective
    return comm.allgather_phase(holo, n_total=n, shape=phases.shape[-2:]), holo
