"""Host-side mirror of the reference modules that call the hot path."""
from .box_regression import Box2BoxTransform
from .caption_consistency import (caption_consistency_loss, caption_consistency_losses,
                                  image_caption_consistency_loss, kd_l1_loss)
from .clip_pretrain import concept_contrastive_loss, image_text_matching_loss, region_concept_distill_loss
from .fast_rcnn import FastRCNNOutputLayers, fast_rcnn_inference, fast_rcnn_inference_single_image
from .gather import GatherLayer
from .matcher import Matcher, pairwise_iou
from .poolers import ROIPooler, convert_boxes_to_pooler_format
from .proposal_utils import decode_proposals, find_top_rpn_proposals, predict_proposals
from .roi_heads import CLIPRes5ROIHeads, ROIHeads, add_ground_truth_to_proposals, label_and_sample_proposals
from .sampling import subsample_labels

__all__ = ["Box2BoxTransform", "caption_consistency_loss", "image_caption_consistency_loss", "caption_consistency_losses", "kd_l1_loss", "region_concept_distill_loss", "concept_contrastive_loss", "image_text_matching_loss",
           "FastRCNNOutputLayers", "fast_rcnn_inference", "fast_rcnn_inference_single_image", "GatherLayer",
           "ROIPooler", "convert_boxes_to_pooler_format", "find_top_rpn_proposals", "predict_proposals", "decode_proposals", "Matcher", "pairwise_iou", "ROIHeads", "CLIPRes5ROIHeads", "label_and_sample_proposals",
           "add_ground_truth_to_proposals", "subsample_labels"]
