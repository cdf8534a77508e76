"""Scene fitting of generated bodies -- drop-in for the reference's FittingOP.

Mirrors source/fitting_habitat.py:42-214 (the self-consistent variant; the
PROX-E script source/fitting_proxe.py:42-214 differs only in the robustifier constant 0.01,
`robust_c`): same constructor dictionaries, `cal_loss(xhr, cam_ext)` returning the four
weighted terms, `fitting(...)` returning the fitted [B,72] vector, `save_result(...)` writing
the reference's per-body pickle.

What changed underneath (none of it changes B=1 results beyond float rounding):
  * SMPL-X, verts_transform, the SDF lookup and the contact NN run in libpsi_b200 kernels;
  * the scene grid and scene points are stored once, not once per body
    (fitting_proxe.py:90,96);
  * contact ids are read once, not 7 JSON files per iteration (cvae.py:99-115, T10);
  * the `.item()` host sync of the collision term (fitting_habitat.py:155) is replaced by a
    branch-free device expression with the same value;
  * only the body->scene direction of the Chamfer term is computed (the loss discards the
    other one, fitting_habitat.py:138);
  * batches: `loss_mode='independent'` (default) sums the reference's B=1 loss over bodies, so
    every body is optimised exactly as the shipped scripts do (batch_size 1,
    fitting_habitat.py:254) and results do not depend on how bodies are sharded;
    `loss_mode='batch'` reproduces the demo notebook's batch-coupled means (demo.ipynb cell 16,
    SURVEY.md T9); both run on the fused engine;
  * the whole iteration (loss, backward, Adam) is captured in a CUDA graph and replayed.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from . import body_model as smplx_b200
from . import chamfer, sdf as sdf_mod
from .geometry import BodyParamParser, GeometryTransformer, VPoserDecoder
from .io import read_scene_vertices


class _Adam:
    """torch.optim.Adam (defaults: betas .9/.999, eps 1e-8) on one tensor with the step count
    on the device, so that an iteration can be replayed from a CUDA graph."""

    def __init__(self, param, lr, betas=(0.9, 0.999), eps=1e-8):
        self.p, self.lr, self.b1, self.b2, self.eps = param, lr, betas[0], betas[1], eps
        self.m = torch.zeros_like(param)
        self.v = torch.zeros_like(param)
        self.t = torch.zeros((), dtype=torch.float64, device=param.device)

    def reset(self):
        self.m.zero_()
        self.v.zero_()
        self.t.zero_()

    @torch.no_grad()
    def step(self, grad):
        # same operation order as torch's single-tensor Adam; bias corrections in float64
        self.t += 1
        self.m.mul_(self.b1).add_(grad, alpha=1 - self.b1)
        self.v.mul_(self.b2).addcmul_(grad, grad, value=1 - self.b2)
        bc1 = 1 - torch.pow(self.b1, self.t)
        bc2_sqrt = (1 - torch.pow(self.b2, self.t)).sqrt()
        step_size = (self.lr / bc1).to(torch.float32)
        denom = (self.v.sqrt() / bc2_sqrt.to(torch.float32)).add_(self.eps)
        self.p.sub_((self.m / denom) * step_size)


class FittingOP:
    def __init__(self, fittingconfig, lossconfig):
        for key, val in fittingconfig.items():
            setattr(self, key, val)
        for key, val in lossconfig.items():
            setattr(self, key, val)
        self.device = torch.device(getattr(self, "device", "cuda"))
        if self.device.type != "cuda":
            raise RuntimeError("FittingOP runs on a CUDA device only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.verbose = getattr(self, "verbose", False)
        self.robust_c = float(getattr(self, "robust_c", 1.0))          # 1.0 habitat/train, 0.01 proxe
        self.loss_mode = getattr(self, "loss_mode", "independent")
        self.use_cuda_graph = bool(getattr(self, "use_cuda_graph", True))
        B = self.batch_size

        # --- VPoser decoder (fitting_habitat.py:54): a real checkpoint dir or synthetic weights
        vw = getattr(self, "vposer_weights", None)
        if vw is not None:
            self.vposer = VPoserDecoder.from_weights(vw)
        elif getattr(self, "vposer_ckpt_path", None):
            # the reference's key: load_vposer(vposer_ckpt_path, vp_model='snapshot'), fitting_habitat.py:54
            self.vposer = VPoserDecoder.from_checkpoint_dir(self.vposer_ckpt_path)
        else:
            raise FileNotFoundError("fittingconfig needs 'vposer_ckpt_path' (a VPoser experiment directory with "
                                    "snapshots/*.pt) or 'vposer_weights' (e.g. synthetic.make_vposer_weights())")
        self.vposer = self.vposer.to(self.device)
        for p_ in self.vposer.parameters():
            p_.requires_grad_(False)

        # --- body model (fitting_habitat.py:57-73)
        # (fittingconfig['body_mesh_model']: an already created module, so that operators for several scenes
        # share ONE device copy of the model constants instead of re-laying the 2 x 64 MB bases out per scene)
        shared_model = getattr(self, "body_mesh_model", None)
        self.body_mesh_model = (shared_model if shared_model is not None else smplx_b200.create(
            getattr(self, "human_model_path", None), model_type="smplx", gender="neutral", ext="npz",
            num_pca_comps=12, create_global_orient=True, create_body_pose=True, create_betas=True,
            create_left_hand_pose=True, create_right_hand_pose=True, create_expression=True,
            create_jaw_pose=True, create_leye_pose=True, create_reye_pose=True, create_transl=True,
            batch_size=B, model_data=getattr(self, "model_data", None))).to(self.device)
        for p_ in self.body_mesh_model.parameters():
            p_.requires_grad_(False)

        self.xhr_rec = torch.randn(B, 75, device=self.device).requires_grad_(True)
        self.optimizer = _Adam(self.xhr_rec, lr=self.init_lr_h)
        # 'adam': the fitting scripts' optimiser (fitting_habitat.py:76).  'lbfgs': the second mode of
        # SURVEY.md T7 / 8(d) -- L-BFGS, history 100, strong-Wolfe line search as
        # human_body_prior/optimizers/lbfgs_ls.py:214-222.  Fused engine: one independent optimiser per body inside
        # libpsi_b200 (csrc/fit_lbfgs.cuh), num_iter = closure evaluations.  Autograd engine: torch.optim.LBFGS
        # (the merged form of the same PR) over the whole batch through the closure, num_iter = max_iter, closure
        # evaluations reported in self.closure_evals -- it couples the bodies through one curvature history.
        self.opt_name = getattr(self, "optimizer_name", "adam")
        if self.opt_name not in ("adam", "lbfgs"):
            raise ValueError("fittingconfig['optimizer_name'] must be 'adam' or 'lbfgs'")
        self.closure_evals = 0

        # --- scene sdf + vertices (fitting_habitat.py:80-96): ONE copy, shared by the batch
        scene = getattr(self, "scene", None)
        if scene is not None:
            self.scene_sdf = sdf_mod.SceneSDF(scene.sdf, scene.grid_min, scene.grid_max, device=self.device)
            pts = scene.points
        else:
            self.scene_sdf = sdf_mod.SceneSDF.from_files(self.scene_sdf_path, device=self.device)
            pts = read_scene_vertices(self.scene_verts_path)
        self.s_verts = torch.tensor(np.asarray(pts), dtype=torch.float32, device=self.device).contiguous()
        # 'index': exact cluster-pruned NN over the static scene (same outputs as 'bruteforce')
        self.nn_mode = getattr(self, "nn", "index")
        if self.nn_mode not in ("index", "bruteforce"):
            raise ValueError("fittingconfig['nn'] must be 'index' or 'bruteforce'")
        self.s_index = chamfer.SceneIndex(self.s_verts) if self.nn_mode == "index" else None

        # --- contact vertex ids, read ONCE (cvae.py:99-115)
        cid = getattr(self, "contact_ids", None)
        if cid is None:
            cid, _ = GeometryTransformer.get_contact_id(self.contact_id_folder, self.contact_part)
        self.contact_ids = torch.as_tensor(np.asarray(cid), dtype=torch.long, device=self.device)
        self._full_contact = bool(self.contact_ids.numel() == self.body_mesh_model._model_data["v_template"].shape[0] and
                                  torch.equal(self.contact_ids, torch.arange(self.contact_ids.numel(), device=self.device)))
        # 'fused': the whole iteration in libpsi_b200 (psi_fit_run, 13 launches/iteration);
        # 'autograd': torch autograd over the psi ops (any loss_mode, any nn mode)
        if self.loss_mode not in ("independent", "batch"):
            raise ValueError("fittingconfig['loss_mode'] must be 'independent' or 'batch'")
        # fused: Adam in either loss mode, or the per-body L-BFGS (independent losses: every body runs its own
        # line search); the batch-coupled L-BFGS of torch.optim (one history for the whole batch) stays on autograd
        self.engine = getattr(self, "engine", "fused" if (self.nn_mode == "index" and (
            self.opt_name == "adam" or self.loss_mode == "independent")) else "autograd")
        if self.opt_name == "lbfgs" and self.engine == "fused" and self.loss_mode != "independent":
            raise ValueError("optimizer_name='lbfgs' with loss_mode='batch' runs on engine='autograd'")
        if self.engine not in ("fused", "autograd"):
            raise ValueError("fittingconfig['engine'] must be 'fused' or 'autograd'")
        self._fused = None
        if self.engine == "fused":
            if self.nn_mode != "index":
                raise ValueError("engine='fused' needs nn='index'")
            from .fused import FusedFit
            lossw = {k: getattr(self, k) for k in ("weight_loss_rec", "weight_loss_vposer", "weight_contact",
                                                   "weight_collision")}
            self._fused = FusedFit(B, self.device, self.body_mesh_model,
                                   self.s_index, self.scene_sdf, self.vposer, self.contact_ids.cpu().numpy(),
                                   lossw, self.robust_c, self.init_lr_h, use_graph=self.use_cuda_graph,
                                   num_streams=getattr(self, "num_streams", None), loss_mode=self.loss_mode,
                                   loop_mode=getattr(self, "loop_mode", None), optimizer=self.opt_name,
                                   lbfgs=dict(lr=float(getattr(self, "lbfgs_lr", 1.0)),
                                              tolerance_grad=float(getattr(self, "tolerance_grad", 1e-5)),
                                              tolerance_change=float(getattr(self, "tolerance_change", 1e-9)),
                                              history_size=int(getattr(self, "history_size", 100))))
        self._graph = None
        self._static_xhr = None
        self._static_cam = None
        self.last_losses = None

    # ------------------------------------------------------------------------------------ loss
    def body_verts(self, xh_rec, cam_ext):
        """xh_rec [B,72] (axis-angle global orient) -> body vertices in the scene frame."""
        bp = BodyParamParser.body_params_encapsulate_batch(xh_rec)
        B = xh_rec.shape[0]
        joint_rot = self.vposer.decode(bp["body_pose_vp"], output_type="aa").view(B, -1)
        out = self.body_mesh_model(return_verts=True, body_pose=joint_rot, transl=bp["transl"],
                                   global_orient=bp["global_orient"], betas=bp["betas"],
                                   left_hand_pose=bp["left_hand_pose"],
                                   right_hand_pose=bp["right_hand_pose"], cam_ext=cam_ext)
        return out.vertices

    def cal_loss(self, xhr, cam_ext):
        """fitting_habitat.py:103-164.  Returns (loss_rec, loss_vposer, loss_contact,
        loss_collision); in 'independent' mode each is the SUM over bodies of the B=1 term."""
        B = self.xhr_rec.shape[0]
        indep = self.loss_mode == "independent"
        red = (lambda t: t.mean(dim=tuple(range(1, t.dim()))).sum()) if indep else (lambda t: t.mean())

        loss_rec = self.weight_loss_rec * red((xhr - self.xhr_rec).abs())
        xh_rec = GeometryTransformer.convert_to_3D_rot(self.xhr_rec)
        loss_vposer = self.weight_loss_vposer * red(xh_rec[:, 16:48] ** 2)

        body_verts = self.body_verts(xh_rec, cam_ext)                        # [B,V,3], scene frame
        contact = body_verts if self._full_contact else body_verts[:, self.contact_ids, :]
        contact_dist, _ = chamfer.nn_distance(contact.contiguous(),
                                              self.s_index if self.s_index is not None else self.s_verts)
        s = torch.sqrt(contact_dist + 1e-4)
        loss_contact = self.weight_contact * red(s / (s + self.robust_c))

        body_sdf = self.scene_sdf.lookup(body_verts)                          # [B,V]
        neg = body_sdf < 0
        if indep:
            cnt = neg.sum(dim=1).clamp(min=1).to(body_sdf.dtype)
            pene = ((-body_sdf) * neg).sum(dim=1).div(cnt).sum()
        else:
            cnt = neg.sum().clamp(min=1).to(body_sdf.dtype)
            pene = ((-body_sdf) * neg).sum() / cnt
        loss_collision = self.weight_collision * pene
        return loss_rec, loss_vposer, loss_contact, loss_collision

    # ------------------------------------------------------------------------------------ loop
    def _iteration(self, xhr, cam):
        self.xhr_rec.grad = None
        lr_, lv, lc, lcol = self.cal_loss(xhr, cam)
        loss = lr_ + lv + lc + lcol
        (grad,) = torch.autograd.grad(loss, self.xhr_rec)
        self.optimizer.step(grad)
        return torch.stack([lr_.detach(), lv.detach(), lc.detach(), lcol.detach()])

    def _build_graph(self):
        self._static_xhr = torch.zeros_like(self.xhr_rec.data)
        self._static_cam = torch.zeros(self._cam_shape, dtype=torch.float32, device=self.device)
        self._static_xhr.copy_(self._xhr_init)
        self._static_cam.copy_(self._cam_init)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._iteration(self._static_xhr, self._static_cam)
        torch.cuda.current_stream(self.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._static_losses = self._iteration(self._static_xhr, self._static_cam)
        self._graph = g

    def fit(self, xh, cam_ext, num_iter=None):
        """xh [B,72] (device), cam_ext [B|1,4,4] camera->scene transform applied to the body
        vertices (already including the y/z flip of fitting_habitat.py:179-184 if wanted).
        Runs `num_iter` Adam iterations, returns the fitted [B,72] (device)."""
        num_iter = self.num_iter if num_iter is None else num_iter
        with torch.cuda.device(self.device):
            xhr = GeometryTransformer.convert_to_6D_rot(xh)
            self._xhr_init, self._cam_init, self._cam_shape = xhr, cam_ext, tuple(cam_ext.shape)
            if self._fused is not None:
                fitted, losses = self._fused.run(xhr, cam_ext, num_iter)
                with torch.no_grad():
                    self.xhr_rec.copy_(fitted)
                self.last_losses = losses.sum(dim=0)
                return GeometryTransformer.convert_to_3D_rot(fitted)
            if self.opt_name == "lbfgs":
                with torch.no_grad():
                    self.xhr_rec.copy_(xhr)
                opt = torch.optim.LBFGS([self.xhr_rec], lr=float(getattr(self, "lbfgs_lr", 1.0)), max_iter=int(num_iter),
                                        max_eval=int(num_iter) * 5 // 4, tolerance_grad=float(getattr(self, "tolerance_grad", 1e-7)),
                                        tolerance_change=float(getattr(self, "tolerance_change", 1e-9)), history_size=100,
                                        line_search_fn="strong_wolfe")
                self.closure_evals = 0

                def closure():
                    opt.zero_grad()
                    terms = self.cal_loss(xhr, cam_ext)
                    loss = terms[0] + terms[1] + terms[2] + terms[3]
                    loss.backward()
                    self.last_losses = torch.stack([t.detach() for t in terms])
                    self.closure_evals += 1
                    return loss

                opt.step(closure)
                return GeometryTransformer.convert_to_3D_rot(self.xhr_rec.detach())
            if self.use_cuda_graph:
                if self._graph is None or self._static_cam.shape != cam_ext.shape:
                    self._build_graph()
                self._static_xhr.copy_(xhr)
                self._static_cam.copy_(cam_ext)
                with torch.no_grad():
                    self.xhr_rec.copy_(xhr)
                self.optimizer.reset()
                for _ in range(num_iter):
                    self._graph.replay()
                self.last_losses = self._static_losses
            else:
                with torch.no_grad():
                    self.xhr_rec.copy_(xhr)
                self.optimizer.reset()
                for ii in range(num_iter):
                    self.last_losses = self._iteration(xhr, cam_ext)
                    if self.verbose:
                        l = self.last_losses.tolist()
                        print("[INFO][fitting] iter={:d}, l_rec={:f}, l_vposer={:f}, l_contact={:f}, "
                              "l_collision={:f}".format(ii, *l))
            return GeometryTransformer.convert_to_3D_rot(self.xhr_rec.detach())

    def connect_batch_shards(self, batch_total, group=None):
        """loss_mode='batch' with the batch sharded over the ranks of a torch.distributed group (one node, one process
        per GPU; SURVEY.md 8(e) option 2): after this call `fit` on every rank optimises the loss of the UNION batch --
        the per-iteration exchange of the penetration counts runs inside the loop through peer memory."""
        if self._fused is None:
            raise _lib.PsiError("connect_batch_shards needs engine='fused'")
        self._fused.connect_group(batch_total, group)

    def trace(self, what):
        """Fused engine only: a buffer of the last iteration the loop evaluated (FusedFit.trace / psi_fit_trace)."""
        if self._fused is None:
            raise _lib.PsiError("trace needs engine='fused'")
        return self._fused.trace(what)

    def profile_iteration(self, xh, cam_ext, warm_iters=20, timed_iters=50):
        """Fused engine only: [(kernel name, mean ms per launch)] of one fitting iteration, each
        launch bracketed by CUDA events on its own stream (psi_fit_profile)."""
        if self._fused is None:
            raise _lib.PsiError("profile_iteration needs engine='fused'")
        with torch.cuda.device(self.device):
            return self._fused.profile(GeometryTransformer.convert_to_6D_rot(xh), cam_ext, warm_iters, timed_iters)

    def fit_host(self, xh_host, cam_ext_host, num_iter=None):
        """End-to-end call with HOST buffers (pinned tensors recommended): H2D of the body
        vectors and camera transform, the fitting loop, D2H of the fitted vectors."""
        xh = xh_host.to(self.device, non_blocking=True)
        cam = cam_ext_host.to(self.device, non_blocking=True)
        out = self.fit(xh, cam, num_iter)
        res = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        res.copy_(out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return res

    def fitting(self, input_data_file):
        """fitting_habitat.py:169-197: one generated body (pickle path or the dict itself)."""
        if isinstance(input_data_file, (str, os.PathLike)):
            with open(input_data_file, "rb") as f:
                body_param_input = pickle.load(f)
        else:
            body_param_input = input_data_file
        xh, self.cam_ext, self.cam_int = BodyParamParser.body_params_parse_fitting(body_param_input, self.device)
        T_mat = torch.eye(4, dtype=torch.float32, device=self.device)
        T_mat[1, 1] = -1.0
        T_mat[2, 2] = -1.0
        flip = bool(getattr(self, "habitat_axis_flip", True))    # fitting_habitat.py:179-184
        trans = torch.matmul(self.cam_ext[:1], T_mat.unsqueeze(0)) if flip else self.cam_ext[:1]
        xh_rec = self.fit(xh, trans)
        print("[INFO][fitting] fitting finish, returning optimal value")
        return xh_rec

    def save_result(self, xh_rec, output_data_file):
        """fitting_habitat.py:201-218 (one pickle per body; as there, a batch overwrites)."""
        dirname = os.path.dirname(output_data_file)
        if dirname and not os.path.exists(dirname):
            os.makedirs(dirname)
        print("[INFO] save results to: " + output_data_file)
        for body_param in BodyParamParser.body_params_encapsulate(xh_rec):
            body_param["cam_ext"] = self.cam_ext.detach().cpu().numpy()
            body_param["cam_int"] = self.cam_int.detach().cpu().numpy()
            with open(output_data_file, "wb") as outfile:
                pickle.dump(body_param, outfile)
