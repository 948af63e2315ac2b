"""ORACLE helper: small non-default constructor configurations of PoseHigherResolutionNet
(pose_higher_hrnet.py:266-287) used to pin the options the W48 teacher does not exercise:
BOTTLENECK stages, 3x3 final convs, several deconv stages / kernel sizes / concat flags."""

VARIANTS = {
    # two 4x4 deconv stages, both concatenating the previous head's output; 3x3 heads
    "two_deconvs_k3_heads": dict(
        final_conv_ksize=3, deconvs=2, deconv_chans=[32, 16], deconv_ksize=[4, 4],
        deconv_num_blocks=1, deconv_cat=[True, True], with_ae_loss=(True, False, False),
        s2_chans=[16, 32], s3_chans=[16, 32, 64], s4_chans=[16, 32, 64, 128],
        s2_blocks=[1, 1], s3_blocks=[1, 1, 1], s4_blocks=[1, 1, 1, 1],
        s3_modules=1, s4_modules=1),
    # BOTTLENECK blocks in stage 2 and 4 (expansion 4: widths 64 / 128 / 256 / 512, which stage 3's BASIC
    # blocks must keep -- the reference applies a non-NoOp transition of an existing branch to the
    # LAST branch, :648-669, so widths cannot change between stages), deconv kernels 3 and 2, no
    # concat for the second deconv
    "bottleneck_stages_k3_k2_deconvs": dict(
        s2_block_type="BOTTLENECK", s4_block_type="BOTTLENECK",
        s2_chans=[16, 32], s3_chans=[64, 128, 256], s4_chans=[16, 32, 64, 128],
        s2_blocks=[1, 2], s3_blocks=[1, 1, 1], s4_blocks=[1, 1, 1, 1],
        s3_modules=1, s4_modules=1,
        deconvs=2, deconv_chans=[32, 32], deconv_ksize=[3, 2], deconv_num_blocks=1,
        deconv_cat=[True, False], with_ae_loss=(True, True, False)),
    # no deconv at all, tag_per_joint off
    "no_deconv_shared_tag": dict(
        tag_per_joint=False, deconvs=0, deconv_chans=[], deconv_ksize=[], deconv_cat=[],
        with_ae_loss=(True,),
        s2_chans=[16, 32], s3_chans=[16, 32, 64], s4_chans=[16, 32, 64, 128],
        s2_blocks=[1, 1], s3_blocks=[1, 1, 1], s4_blocks=[1, 1, 1, 1], s3_modules=1, s4_modules=1),
}


def variant_kwargs(name):
    return {k: (list(v) if isinstance(v, list) else v) for k, v in VARIANTS[name].items()}
