"""mvlt-b200: B200-native (sm_100a) forward hot path of MVLT — Swin-S visual extractor + BERT-base joint image-text
encoder + VQA / retrieval / pretraining heads + N x N retrieval scoring — behind the reference's nn.Module API.

    from medical_vision_langauge_transformer_b200.modules.model import MVLBertForVQA, MVLBertForRetrieval, MVLBertForPretraining
    from medical_vision_langauge_transformer_b200.modules.config import MVLBertConfigforVQA, MVLBertRetrieval, MVLBertPretrainConfig

Importing the package does not touch CUDA; the first op loads libmvlt_b200.so (built by `build.py`) and raises if it
is missing — there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
