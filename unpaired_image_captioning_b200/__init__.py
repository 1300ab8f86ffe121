"""B200-native attention-LSTM caption decoder: drop-in for the hot path of
gujiuxiang/unpaired_image_captioning (pivot_based_eccv2018/models + misc/criterion).

    import unpaired_image_captioning_b200 as uic
    model = uic.setup(opt).cuda()                       # models.setup(opt)
    logprobs = model(fc, attri, att, labels, att_masks) # teacher-forced forward
    seq, lp = model(fc, attri, att, att_masks, opt={'beam_size': 3}, mode='sample')
    loss = uic.LanguageModelCriterion(opt)(logprobs, labels[:, 1:], masks[:, 1:])
"""
from .criterion import LanguageModelCriterion, RewardCriterion  # noqa: F401
from .loader import FeatureCache, FeatureStream, decode_split, gather_captions, shard_bounds  # noqa: F401
from .models import (Att2in2Core, Att2in2Model, AttModel, Attention, CaptionModel, TopDownCore,  # noqa: F401
                     TopDownModel, setup)

__all__ = ["setup", "LanguageModelCriterion", "RewardCriterion", "AttModel", "Att2in2Model", "TopDownModel", "Attention",
           "Att2in2Core", "TopDownCore", "CaptionModel", "FeatureCache", "FeatureStream", "decode_split", "gather_captions",
           "shard_bounds"]
