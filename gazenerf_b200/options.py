"""Host-side mirror of the reference's option object (configs/gazenerf_options.py:1-35).

Same attribute names and defaults, same ``para_dict`` override of featmap_size / featmap_nc / pred_img_size,
so callers that build ``BaseOptions`` for the reference can pass the object unchanged (duck-typed).
"""


class BaseOptions(object):
    def __init__(self, para_dict=None) -> None:
        self.bg_type = "white"
        self.iden_code_dims = 100
        self.expr_code_dims = 79
        self.text_code_dims = 100
        self.illu_code_dims = 27
        self.eye_code_dims = 2
        self.auxi_shape_code_dims = 179
        self.auxi_appea_code_dims = 127
        self.num_sample_coarse = 64
        self.num_sample_fine = 128
        self.world_z1 = 2.5
        self.world_z2 = -3.5
        self.mlp_hidden_nchannels = 384
        if para_dict is None:
            self.featmap_size = 64
            self.featmap_nc = 258
            self.pred_img_size = 512
        else:
            self.featmap_size = para_dict["featmap_size"]
            self.featmap_nc = para_dict["featmap_nc"]
            self.pred_img_size = para_dict["pred_img_size"]
