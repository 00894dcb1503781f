"""TEST INFRASTRUCTURE -- the reference CALLER of the hot path, restated statement by statement.

`BetaModelCaller` holds the seven parameter tensors the way `BetaModel` does and its `render()` body is the body of
`BetaModel.render` (scene/beta_model.py:660-722); the getters are those of scene/beta_model.py:103-159.  scene/ itself
cannot be imported (plyfile, sklearn-based init, viser, fused_ssim are absent) and /root/reference does not
exist on the GPU box, so the statements are restated here; tests/test_abi_and_host.py compares the AST of every
restated method with the AST `ast` extracts from the reference file whenever /root/reference is present, so "the
literal statements" is checked, not claimed.

It imports `gsplat` the way the reference does.  With tests/conftest.py's sys.path that is the ubs_b200 shim.
"""
import math

import torch
import torch.nn.functional as F
from gsplat.cuda._wrapper import cond_mean_convariance_opacity, l_triangle_to_rotmat, rot_scale_l_triangle_to_covar
from gsplat.rendering import rasterization


class ViewpointCamera:
    """The attributes of scene/cameras.py:18-95 that BetaModel.render reads."""

    def __init__(self, cam):
        # cam: ubs_b200.synth.Camera
        self.image_width, self.image_height = cam.width, cam.height
        self.FoVx = 2.0 * math.atan(cam.width / (2.0 * float(cam.K[0, 0])))
        self.FoVy = 2.0 * math.atan(cam.height / (2.0 * float(cam.K[1, 1])))
        self.world_view_transform = cam.viewmat.transpose(0, 1).contiguous()  # stored transposed (cameras.py:69-73)
        self.projection_matrix = torch.zeros(4, 4, device=cam.viewmat.device)  # only its .device is read
        self.camera_center = cam.cam_pos
        self.timestamp = cam.timestamp

    def K(self):
        """The intrinsics BetaModel.render builds from the fields above (beta_model.py:664-673)."""
        K = torch.zeros((3, 3), device=self.projection_matrix.device)
        K[0, 0] = 0.5 * self.image_width / math.tan(self.FoVx / 2)
        K[1, 1] = 0.5 * self.image_height / math.tan(self.FoVy / 2)
        K[0, 2] = self.image_width / 2
        K[1, 2] = self.image_height / 2
        K[2, 2] = 1.0
        return K


class BetaModelCaller:
    # names of the restated methods whose AST is compared with the reference's (tests/test_abi_and_host.py)
    RESTATED = ("setup_functions", "get_scale", "get_l_triangle", "get_mean", "get_opacity", "get_beta", "get_rotation",
                "get_covariance", "get_xyz_covariance", "get_cond_mean_convariance_opacity", "render")

    def setup_functions(self):
        def beta_activation(betas):
            return 4.0 * torch.exp(betas)

        def inverse_softplus(y):
            return y + torch.log(-torch.expm1(-y))

        self.scale_activation = F.softplus
        self.scale_inverse_activation = inverse_softplus

        self.opacity_activation = torch.sigmoid
        self.inverse_opacity_activation = inverse_sigmoid

        self.beta_activation = beta_activation

        self.l_triangs_activation = lambda x: x
        self.l_triangs_inverse_activation = lambda x: x

    def __init__(self, scene, background, requires_grad=False):
        self.input_dim = scene.D

        def leaf(t):
            return t.detach().clone().requires_grad_(requires_grad)

        self._xyz, self._mean, self._rgb = leaf(scene.xyz), leaf(scene.mean), leaf(scene.rgb)
        self._opacity, self._beta = leaf(scene.opacity), leaf(scene.beta)
        self._scale, self._l_triangle = leaf(scene.scale), leaf(scene.l_triangle)
        self.background = background
        self.setup_functions()
        # scene/beta_model.py:69-73
        tril_i, tril_j = torch.tril_indices(self.input_dim, self.input_dim, offset=-1)
        mask_rest = (tril_i >= 3) | (tril_j >= 3)
        self.rest_i = tril_i[mask_rest].to(torch.int32).to("cuda")
        self.rest_j = tril_j[mask_rest].to(torch.int32).to("cuda")

    def leaves(self):
        return [self._xyz, self._mean, self._rgb, self._opacity, self._beta, self._scale, self._l_triangle]

    @property
    def get_scale(self):
        return self.scale_activation(self._scale)

    @property
    def get_l_triangle(self):
        return self.l_triangs_activation(self._l_triangle)

    @property
    def get_mean(self):
        return torch.cat([self._xyz, self._mean], dim=-1)

    @property
    def get_opacity(self):
        return self.opacity_activation(self._opacity)

    @property
    def get_beta(self):
        return self.beta_activation(self._beta)

    @property
    def get_rotation(self):
        return l_triangle_to_rotmat(self.get_l_triangle[:, :3])

    @property
    def get_covariance(self):
        return rot_scale_l_triangle_to_covar(
            self.get_rotation,
            self.get_scale,
            self.get_l_triangle,
            self.rest_i,
            self.rest_j,
        )

    @property
    def get_xyz_covariance(self):
        return rot_scale_l_triangle_to_covar(
            self.get_rotation,
            self.get_scale,
            self.get_l_triangle,
            self.rest_i,
            self.rest_j,
            spatial_block=True,
        )

    def get_cond_mean_convariance_opacity(self, q):
        v = self.get_covariance
        m = self.get_mean
        o = self.get_opacity
        b = self.get_beta[:, 1:]
        return cond_mean_convariance_opacity(m, v, o, b, q)

    def render(self, viewpoint_camera, render_mode="RGB", mask=None):
        if mask == None:
            mask = torch.ones_like(self.get_opacity.squeeze()).bool()

        K = torch.zeros((3, 3), device=viewpoint_camera.projection_matrix.device)

        fx = 0.5 * viewpoint_camera.image_width / math.tan(viewpoint_camera.FoVx / 2)
        fy = 0.5 * viewpoint_camera.image_height / math.tan(viewpoint_camera.FoVy / 2)

        K[0, 0] = fx
        K[1, 1] = fy
        K[0, 2] = viewpoint_camera.image_width / 2
        K[1, 2] = viewpoint_camera.image_height / 2
        K[2, 2] = 1.0

        if self.input_dim > 3:
            cam_pos = viewpoint_camera.camera_center
            view_dir = self._xyz - cam_pos.unsqueeze(0)
            view_dir = view_dir / view_dir.norm(dim=-1, keepdim=True)
            if self.input_dim == 6:
                query = view_dir
            elif self.input_dim == 7:
                timestamp = torch.full(
                    (view_dir.shape[0], 1),
                    viewpoint_camera.timestamp,
                    device=view_dir.device,
                    dtype=view_dir.dtype,
                )
                query = torch.cat([view_dir, timestamp], dim=-1)
            else:
                raise NotImplementedError("Only implemented for 6D or 7D query")
            means, convs, opacities = self.get_cond_mean_convariance_opacity(query)
        else:
            means = self.get_mean
            convs = self.get_covariance
            opacities = self.get_opacity

        rgbs, alphas, meta = rasterization(
            means=means[mask],
            l_triagnles=self.get_l_triangle[mask],
            scales=self.get_scale[mask],
            opacities=opacities.squeeze()[mask],
            betas=self.get_beta[:, :1].squeeze()[mask],
            colors=self._rgb[mask],
            viewmats=viewpoint_camera.world_view_transform.transpose(0, 1).unsqueeze(0),
            Ks=K.unsqueeze(0),
            width=viewpoint_camera.image_width,
            height=viewpoint_camera.image_height,
            backgrounds=self.background.unsqueeze(0),
            render_mode=render_mode,
            covars=convs[mask],
        )

        # # Convert from N,H,W,C to N,C,H,W format
        rgbs = rgbs.permute(0, 3, 1, 2).contiguous()[0]

        return {
            "render": rgbs,
            "viewspace_points": meta["means2d"],
            "visibility_filter": meta["radii"] > 0,
            "radii": meta["radii"],
            "is_used": meta["radii"] > 0,
        }

    # ---- the rasterization() call of BetaModel.view (scene/beta_model.py:724-831) without the GUI state object ----
    @torch.no_grad()
    def view_call(self, c2w, K, W, H, query, mask, render_mode, near_plane, far_plane, radius_clip):
        means, convs, opacities = self.get_cond_mean_convariance_opacity(query)
        render_colors, alphas, meta = rasterization(
            means=means[mask],
            l_triagnles=self.get_l_triangle[mask],
            scales=self.get_scale[mask],
            opacities=opacities.squeeze()[mask],
            betas=self.get_beta[:, :1].squeeze()[mask],
            colors=self._rgb[mask],
            viewmats=torch.linalg.inv(c2w).unsqueeze(0),
            Ks=K.unsqueeze(0),
            width=W,
            height=H,
            backgrounds=self.background.unsqueeze(0),
            render_mode=render_mode if render_mode != "Alpha" else "RGB",
            covars=convs[mask],
            near_plane=near_plane,
            far_plane=far_plane,
            radius_clip=radius_clip,
        )
        rendered_count_number = (meta["radii"] > 0).sum().item()
        if render_mode == "Alpha":
            render_colors = alphas
        return render_colors, rendered_count_number


def inverse_sigmoid(x):  # utils/general_utils.py (imported by scene/beta_model.py:3)
    return torch.log(x / (1 - x))
