"""The state_dict contract of the DANBO field: names, shapes, order.

Names are the reference's own (core/networks/danbo.py:104-185, nerf.py:73-105, gnn_backbone.py:184-222,
608-629,651-681, misc.py:129-156, embedding.py:4-15) so that reference checkpoints load unchanged.
ParallelLinear weights are (joint, in, out); nn.Linear weights are (out, in).
"""
from collections import OrderedDict

J = 24


def danbo_param_shapes(n_framecodes=8, W=256, view_W=128, node_W=128, agg_W=32, voxel_feat=5, voxel_res=16,
                       multires_voxel=6, multires_graph=5, multires_views=4, framecode_ch=128, opt_framecode=True):
    """opt_framecode=False (configs/surreal/danbo_*.txt): no per-frame code - the view layer takes 256 + 27 inputs and
    `framecodes.codes.weight` does not exist."""
    if not opt_framecode:
        framecode_ch = 0
    feat = voxel_feat * 3                       # FGNNcat: three axis lines concatenated
    x_ch = feat * (1 + 2 * multires_voxel)      # 195
    g_ch = 6 * (1 + 2 * multires_graph)         # 66
    v_ch = 3 * (1 + 2 * multires_views)         # 27
    s = OrderedDict()
    s["pts_linears.0.weight"] = (W, x_ch)
    s["pts_linears.0.bias"] = (W,)
    for i in range(1, 8):
        s[f"pts_linears.{i}.weight"] = (W, W + x_ch) if i == 5 else (W, W)
        s[f"pts_linears.{i}.bias"] = (W,)
    s["alpha_linear.weight"] = (1, W)
    s["alpha_linear.bias"] = (1,)
    s["views_linears.0.weight"] = (view_W, v_ch + framecode_ch + 2 * view_W)
    s["views_linears.0.bias"] = (view_W,)
    s["feature_linear.weight"] = (2 * view_W, W)
    s["feature_linear.bias"] = (2 * view_W,)
    s["rgb_linear.weight"] = (3, view_W)
    s["rgb_linear.bias"] = (3,)
    if opt_framecode:
        s["framecodes.codes.weight"] = (n_framecodes, framecode_ch)
    s["graph_net.axis_scale"] = (J, 3)
    for i, cin in ((0, g_ch), (1, node_W)):
        s[f"graph_net.layers.{i}.bias"] = (node_W,)
        s[f"graph_net.layers.{i}.adj_w"] = (1, J, J)
        s[f"graph_net.layers.{i}.adj"] = (1, J, J)          # buffer
        s[f"graph_net.layers.{i}.lin.weight"] = (J, cin, node_W)
    s["graph_net.layers.2.weight"] = (J, node_W, node_W)
    s["graph_net.layers.2.bias"] = (1, J, node_W)
    s["graph_net.layers.3.weight"] = (J, node_W, voxel_feat * voxel_res * 3)
    s["graph_net.layers.3.bias"] = (1, J, voxel_feat * voxel_res * 3)
    s["prob_linears.layers.0.bias"] = (agg_W,)
    s["prob_linears.layers.0.adj_w"] = (1, J, J)
    s["prob_linears.layers.0.adj"] = (1, J, J)              # buffer
    s["prob_linears.layers.0.lin.weight"] = (J, feat, agg_W)
    s["prob_linears.layers.1.weight"] = (J, agg_W, agg_W)
    s["prob_linears.layers.1.bias"] = (1, J, agg_W)
    s["prob_linears.layers.2.weight"] = (J, agg_W, 1)
    s["prob_linears.layers.2.bias"] = (1, J, 1)
    return s


def anerf_param_shapes(n_framecodes=8, W=448, view_W=224, multires=7, multires_views=4, framecode_ch=128):
    """A-NeRF field (nerf_type=nerf; configs/h36m_zju/anerf_base.txt; nerf.py:73-105): density input 24*(1+2*7) + 72
    = 432, view input 72*(1+2*4) + 128 = 776.  The cutoff embedders' `cutoff_dist` / `tau` entries of the reference
    state_dict (cutoff_embedder.py:128-133) are constants of the path (0.5, 20) and are not listed here."""
    x_ch = J * (1 + 2 * multires) + J * 3
    v_ch = J * 3 * (1 + 2 * multires_views)
    s = OrderedDict()
    s["pts_linears.0.weight"] = (W, x_ch)
    s["pts_linears.0.bias"] = (W,)
    for i in range(1, 8):
        s[f"pts_linears.{i}.weight"] = (W, W + x_ch) if i == 5 else (W, W)
        s[f"pts_linears.{i}.bias"] = (W,)
    s["alpha_linear.weight"] = (1, W)
    s["alpha_linear.bias"] = (1,)
    s["views_linears.0.weight"] = (view_W, v_ch + framecode_ch + 2 * view_W)
    s["views_linears.0.bias"] = (view_W,)
    s["feature_linear.weight"] = (2 * view_W, W)
    s["feature_linear.bias"] = (2 * view_W,)
    s["rgb_linear.weight"] = (3, view_W)
    s["rgb_linear.bias"] = (3,)
    s["framecodes.codes.weight"] = (n_framecodes, framecode_ch)
    return s


BUFFER_NAMES = ("graph_net.layers.0.adj", "graph_net.layers.1.adj", "prob_linears.layers.0.adj")
