"""B200-native (sm_100a) implementation of the TwoTower news-recommendation hot path of
tyh666/News-Recommendation-MIND, drop-in behind the reference's twotower.py / utils/Manager.py.

    from news_recommendation_mind_b200 import TwoTower, BERT_Embedding, CNN_Encoder, RNN_User_Encoder

All arithmetic runs in libmindrec.so (csrc/, C ABI in include/mindrec.h); there is no CPU fallback.
"""
from .modules import (Attention_Pooling, Average_Pooling, BERT_Embedding, CNN_Encoder, LSTUR, LSTUR_User_Encoder,
                      MHA_Encoder, MHA_User_Encoder, MultiheadAttention, RNN_User_Encoder)
from .twotower import TwoTower, TwoTowerBaseModel

__all__ = ["TwoTower", "TwoTowerBaseModel", "BERT_Embedding", "CNN_Encoder", "RNN_User_Encoder", "LSTUR_User_Encoder",
           "LSTUR", "Attention_Pooling", "Average_Pooling", "MHA_Encoder", "MHA_User_Encoder", "MultiheadAttention"]
