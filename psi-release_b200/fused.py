"""The fused fitting loop (psi_fit_* in include/psi_b200.h): one C call runs all iterations of
source/fitting_habitat.py:177-191 for a batch of bodies -- VPoser decode, rotation chain, SMPL-X,
contact NN, SDF, losses, backward and Adam as 13 kernel launches per iteration from a CUDA graph.

The loop runs on a joint-coherent re-ordering of the model's vertices (body_model.SMPLX.coherent_handle): vertices
never leave the library, only the fitted parameters do; `trace()` hands per-vertex buffers back in the file's order.

A batch CAN be split over several contexts (`num_streams` / PSI_FIT_STREAMS), each with its own stream and graph;
results do not depend on the split (every kernel is per-body deterministic), but it is slower (657 vs 761 bodies/s:
the blend GEMMs stream the whole basis per context), so the default is one context.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib


def _kd_runs(pts, ids, leaf=32):
    """Order `ids` so that aligned runs of `leaf` are compact kd cells (median split of the longest
    axis, left part a multiple of `leaf`) -- the same construction as the scene index."""
    out = []
    stack = [np.asarray(ids, dtype=np.int64)]
    while stack:
        cur = stack.pop()
        if len(cur) <= leaf:
            out.append(cur)
            continue
        p = pts[cur]
        ax = int(np.argmax(p.max(0) - p.min(0)))
        nl = ((len(cur) // leaf + 1) // 2) * leaf
        o = np.argsort(p[:, ax], kind="stable")
        stack.append(cur[o[nl:]])          # popped after the left part: keeps left-to-right order
        stack.append(cur[o[:nl]])
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def _spatial_order(ids, v_template, weights=None, parents=None):
    """Order the contact vertex ids so that 32 consecutive ids stay neighbours on the POSED body:
    vertices are grouped by the joint that dominates their skinning weights (they move together),
    joints are laid out depth-first along the kinematic tree (parent and child bones touch), and
    inside a joint the vertices follow a kd order of the template.  Duplicate ids stay adjacent.
    The NN result does not depend on the order; the group schedule's speed does."""
    vt = np.asarray(v_template, dtype=np.float64)
    ids = np.asarray(ids, dtype=np.int64)
    uniq, counts = np.unique(ids, return_counts=True)
    if weights is None or parents is None:
        order = _kd_runs(vt, uniq)
    else:
        owner = np.asarray(weights)[uniq].argmax(1)
        parents = np.asarray(parents, dtype=np.int64)
        nj = len(parents)
        kids = [[] for _ in range(nj)]
        for j in range(1, nj):
            kids[int(parents[j])].append(j)
        dfs, stack = [], [0]
        while stack:
            j = stack.pop()
            dfs.append(j)
            stack.extend(reversed(kids[j]))
        order = np.concatenate([_kd_runs(vt, uniq[owner == j]) for j in dfs] + [np.zeros(0, np.int64)])
    mult = dict(zip(uniq.tolist(), counts.tolist()))
    return np.ascontiguousarray(np.repeat(order, [mult[int(v)] for v in order]).astype(np.int32))


class FusedFit:
    """Owns psi_fit_ctx objects bound to (body model handle, scene index, scene SDF, VPoser decoder)."""

    def __init__(self, batch_size, model_handle, body_model, scene_index, scene_sdf, vposer, contact_ids,
                 weights, robust_c, lr, use_graph=True, num_streams=None, loss_mode="independent", loop_mode=None,
                 optimizer="adam", lbfgs=None):
        if scene_sdf.num_scenes != 1:
            raise ValueError("the fused loop fits one scene per context")
        # model_handle: a body_model._ModelHandle, or just the device (the handle is then made here, once, in the
        # vertex order the loop wants)
        dev_only = not hasattr(model_handle, "h")
        self.device = torch.device(model_handle) if dev_only else model_handle.device
        self.B = int(batch_size)
        # Vertices never leave the loop (only the fitted parameters do), so it runs on the joint-coherent vertex order
        # (body_model.SMPLX.coherent_handle); trace() hands per-vertex buffers back in the file's order.
        # PSI_FIT_VORDER=0 keeps the file's order.
        self._perm = self._inv = None
        md = body_model._model_data
        vt, wts = md["v_template"], md["weights"]
        if os.environ.get("PSI_FIT_VORDER", "1") != "0":
            model_handle = body_model.coherent_handle(self.device)
            perm, inv = model_handle.perm, model_handle.inv
            self._perm = torch.as_tensor(perm, device=self.device)
            self._inv = torch.as_tensor(inv, device=self.device)
            contact_ids = inv[np.asarray(contact_ids, dtype=np.int64)]
            vt, wts = vt[perm], wts[perm]
        elif dev_only:
            model_handle = body_model.handle(self.device)
        self._keep = (model_handle, scene_index, scene_sdf)     # borrowed device objects
        f32 = lambda t: np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))
        sd = {k: f32(v) for k, v in vposer.state_dict().items()}
        W1, b1 = sd["bodyprior_dec_fc1.weight"], sd["bodyprior_dec_fc1.bias"]
        W2, b2 = sd["bodyprior_dec_fc2.weight"], sd["bodyprior_dec_fc2.bias"]
        W3, b3 = sd["bodyprior_dec_out.weight"], sd["bodyprior_dec_out.bias"]
        hl, hr = f32(body_model.left_hand_components), f32(body_model.right_hand_components)
        pm = f32(body_model.pose_mean)
        # Query order = order of first appearance in the id list: spatially sorted ids make the 32
        # queries of a warp neighbours on the body (needed by the thread-per-query NN schedule);
        # the result does not depend on the order.
        cid = _spatial_order(contact_ids, vt, wts, np.asarray(md["kintree_table"])[0])
        if num_streams is None:
            # measured: two half-batch contexts are not faster (kernels do not shrink with B); PSI_FIT_STREAMS re-measures
            num_streams = int(os.environ.get("PSI_FIT_STREAMS", "1"))
        num_streams = max(1, min(int(num_streams), self.B))
        if loss_mode not in ("independent", "batch"):
            raise ValueError("loss_mode must be 'independent' or 'batch'")
        if loss_mode == "batch" and num_streams != 1:
            raise ValueError("loss_mode='batch' couples the bodies of a batch: one context (num_streams=1)")
        # 'whole' = all iterations in ONE graph launch (conditional WHILE node); 'replay' = one launch per iteration
        if loop_mode is None:
            loop_mode = os.environ.get("PSI_FIT_LOOP", "whole")
        if loop_mode not in ("whole", "replay"):
            raise ValueError("loop_mode must be 'whole' or 'replay'")
        self.loss_mode, self.loop_mode = loss_mode, loop_mode
        if optimizer not in ("adam", "lbfgs"):
            raise ValueError("optimizer must be 'adam' or 'lbfgs'")
        if optimizer == "lbfgs" and loss_mode != "independent":
            raise ValueError("the per-body L-BFGS needs loss_mode='independent'")
        self.optimizer = optimizer
        lb = dict(lr=1.0, tolerance_grad=1e-5, tolerance_change=1e-9, history_size=100, zoom_max=300)   # lbfgs_ls.py:214-222
        lb.update(lbfgs or {})
        base, rem = divmod(self.B, num_streams)
        self.parts = []                  # (start, size, handle)
        self.xdim = 19 + W1.shape[1] + 2 * hl.shape[0]
        hp = lambda a: ctypes.c_void_p(a.ctypes.data)
        start = 0
        for i in range(num_streams):
            nb = base + (1 if i < rem else 0)
            cfg = _lib.FitConfig(B=nb, use_graph=1 if use_graph else 0,
                                 w_rec=float(weights["weight_loss_rec"]), w_vposer=float(weights["weight_loss_vposer"]),
                                 w_contact=float(weights["weight_contact"]), w_collision=float(weights["weight_collision"]),
                                 robust_c=float(robust_c), lr=float(lr), beta1=0.9, beta2=0.999, eps=1e-8,
                                 nn_mode=int(os.environ.get("PSI_FIT_NN_MODE", "0")),
                                 loop_mode=0 if loop_mode == "whole" else 1,
                                 loop_unroll=int(os.environ.get("PSI_FIT_UNROLL", "0")),
                                 loss_mode=1 if loss_mode == "batch" else 0, optimizer=1 if optimizer == "lbfgs" else 0,
                                 lbfgs_lr=float(lb["lr"]), lbfgs_tolerance_grad=float(lb["tolerance_grad"]),
                                 lbfgs_tolerance_change=float(lb["tolerance_change"]), lbfgs_history=int(lb["history_size"]),
                                 lbfgs_zoom_max=int(lb["zoom_max"]))
            h = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                rc = _lib.lib().psi_fit_create(
                    ctypes.byref(h), model_handle.h, model_handle.V, model_handle.J, model_handle.NB,
                    scene_index.h, _lib.ptr(scene_index.points), _lib.ptr(scene_sdf.sdf), scene_sdf.dim,
                    hp(scene_sdf.grid_min), hp(scene_sdf.grid_max), hp(W1), hp(b1), hp(W2), hp(b2), hp(W3), hp(b3),
                    W1.shape[1], W1.shape[0], W3.shape[0] // 6, hp(hl), hp(hr), hp(pm), hl.shape[0],
                    hp(cid), cid.shape[0], ctypes.byref(cfg), _lib.stream_ptr())
            _lib.check(rc, "psi_fit_create")
            self.parts.append((start, nb, h))
            start += nb

    # ---- loss_mode='batch' sharded over several contexts / GPUs (psi_fit_set_peers) -----------------------------
    def connect_local(self, shards, rank, batch_total):
        """Shards living in THIS process (a list of FusedFit, this one at index `rank`): device pointers."""
        if self.loss_mode != "batch" or len(self.parts) != 1:
            raise ValueError("sharding the batch-coupled loss needs loss_mode='batch' and one context per shard")
        L = _lib.lib()
        ptrs = (ctypes.c_void_p * len(shards))(*[L.psi_fit_exchange_ptr(s.parts[0][2]) for s in shards])
        with torch.cuda.device(self.device):
            _lib.check(L.psi_fit_set_peers(self.parts[0][2], int(rank), len(shards), int(batch_total), None, ptrs), "psi_fit_set_peers")
        self.world = len(shards)

    def connect_group(self, batch_total, group=None):
        """Shards = the ranks of a torch.distributed group on ONE node (one process per GPU): every rank's exchange
        buffer is opened in every other rank through its CUDA IPC handle; no collective runs afterwards -- each
        iteration's 8 bytes per peer travel as plain stores over the peer mapping inside the loop's graph."""
        import torch.distributed as dist
        if self.loss_mode != "batch" or len(self.parts) != 1:
            raise ValueError("sharding the batch-coupled loss needs loss_mode='batch' and one context per shard")
        L = _lib.lib()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        mine = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(L.psi_fit_exchange_handle(self.parts[0][2], mine), "psi_fit_exchange_handle")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine.raw), group=group)
        bufs = [ctypes.create_string_buffer(h, 64) for h in handles]
        arr = (ctypes.c_void_p * world)(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
        with torch.cuda.device(self.device):
            _lib.check(L.psi_fit_set_peers(self.parts[0][2], rank, world, int(batch_total), arr, None), "psi_fit_set_peers")
        self.world = world
        dist.barrier(group)                       # every rank has mapped every buffer before anyone starts a loop

    def begin(self, xhr, cam_ext, num_iter):
        """Enqueue the loop (it runs on the context's own stream) and return; `end()` collects the result.  Lets
        several shards of one process run concurrently."""
        _lib.require_cuda(xhr, cam_ext)
        self._xhr = xhr.contiguous().float()
        self._cam = cam_ext.reshape(-1, 16)[:, :12].contiguous().float()
        shared = self._cam.shape[0] == 1
        with torch.cuda.device(self.device):
            for s, nb, h in self.parts:
                _lib.check(_lib.lib().psi_fit_begin(h, _lib.ptr(self._xhr[s:s + nb]), _lib.ptr(self._cam if shared else self._cam[s:s + nb]),
                                                    0 if shared else 12, int(num_iter), _lib.stream_ptr()), "psi_fit_begin")

    def end(self):
        out = torch.empty_like(self._xhr)
        losses = torch.empty(self.B, 4, dtype=torch.float32, device=self._xhr.device)
        with torch.cuda.device(self.device):
            for s, nb, h in self.parts:
                _lib.check(_lib.lib().psi_fit_end(h, _lib.ptr(out[s:s + nb]), _lib.ptr(losses[s:s + nb]), _lib.stream_ptr()), "psi_fit_end")
        return out, losses

    def run(self, xhr, cam_ext, num_iter):
        """xhr [B,75] device; cam_ext [B|1,4,4] -> (fitted xhr [B,75], losses [B,4])."""
        _lib.require_cuda(xhr, cam_ext)
        if xhr.shape != (self.B, self.xdim):
            raise ValueError(f"xhr must be [{self.B},{self.xdim}], got {tuple(xhr.shape)}")
        xhr = xhr.contiguous().float()
        cam = cam_ext.reshape(-1, 16)[:, :12].contiguous().float()
        if cam.shape[0] not in (1, self.B):
            raise ValueError("cam_ext must be [B,4,4] or [1,4,4]")
        shared = cam.shape[0] == 1
        out = torch.empty_like(xhr)
        losses = torch.empty(self.B, 4, dtype=torch.float32, device=xhr.device)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            for s, nb, h in self.parts:          # enqueue every part, then join them
                rc = L.psi_fit_begin(h, _lib.ptr(xhr[s:s + nb]), _lib.ptr(cam if shared else cam[s:s + nb]),
                                     0 if shared else 12, int(num_iter), st)
                _lib.check(rc, "psi_fit_begin")
            for s, nb, h in self.parts:
                rc = L.psi_fit_end(h, _lib.ptr(out[s:s + nb]), _lib.ptr(losses[s:s + nb]), st)
                _lib.check(rc, "psi_fit_end")
        return out, losses

    def trace(self, what):
        """One buffer of the most recent iteration the loop evaluated (psi_fit_trace): 'x_eval' [B,75] (where it
        was evaluated), 'grad_x' [B,75] (dL/dx there, before Adam), 'verts' [B,V,3], 'sdf' [B,V], 'sdf_grad'
        [B,V,3], 'nn_dist' / 'nn_idx' [B,nu] in query order, 'query_ids' [nu] (body vertex of each query slot),
        'losses' [B,4], 'x' [B,75] (after the step), 'adam_m' / 'adam_v', 'pose6d' [B,22,6].  Device tensors."""
        code, is_int = _lib.FIT_TRACE[what]
        L = _lib.lib()
        outs = []
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            for _, nb, h in self.parts:
                nbytes = int(L.psi_fit_trace_bytes(h, code))
                t = torch.empty(nbytes // 4, dtype=torch.int32 if is_int else torch.float32, device=self.device)
                _lib.check(L.psi_fit_trace(h, code, _lib.ptr(t), nbytes, st), "psi_fit_trace")
                outs.append(t if what == "query_ids" else t.view(1, -1) if what == "exchange" else t.view(nb, -1))
        if what == "query_ids":
            return outs[0] if self._perm is None else self._perm[outs[0].long()].to(torch.int32)
        out = torch.cat(outs, 0)
        if self._inv is not None and what in ("verts", "sdf", "sdf_grad"):      # back to the file's vertex order
            V = self._inv.numel()
            out = out.view(out.shape[0], V, -1)[:, self._inv].reshape(out.shape[0], -1)
        return out

    def profile(self, xhr, cam_ext, warm_iters=20, timed_iters=50):
        """Per-kernel timing of one fitting iteration (psi_fit_profile: eager launches with a CUDA
        event behind each, on the launching stream) -> list of (kernel name, mean ms per launch)."""
        _lib.require_cuda(xhr, cam_ext)
        xhr = xhr.contiguous().float()
        cam = cam_ext.reshape(-1, 16)[:, :12].contiguous().float()
        shared = cam.shape[0] == 1
        s, nb, h = self.parts[0]
        ms = (ctypes.c_float * 64)()
        names = (ctypes.c_char_p * 64)()
        with torch.cuda.device(self.device):
            n = _lib.lib().psi_fit_profile(h, _lib.ptr(xhr[s:s + nb]), _lib.ptr(cam if shared else cam[s:s + nb]),
                                           0 if shared else 12, int(warm_iters), int(timed_iters), ms, names, 64,
                                           _lib.stream_ptr())
        if n < 0:
            _lib.check(-n - 1000, "psi_fit_profile")
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def __del__(self):
        try:
            for _, _, h in getattr(self, "parts", []):
                _lib.lib().psi_fit_destroy(h)
            self.parts = []
        except Exception:
            pass
