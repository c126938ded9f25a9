"""text_alignment_b200 -- B200-native implementation of DDMAL/text_alignment's one
data-parallel hot path: the affine-gap Needleman-Wunsch of textSeqCompare.perform_alignment.

    from text_alignment_b200 import textSeqCompare as tsc      # drop-in module
    tra_align, ocr_align = tsc.perform_alignment(list(transcript), list(ocr))
"""
__version__ = '0.1.0'
