"""Drop-in import surface of Suzie14/Bayesformer (`from src.model.transformer import MyTransformer`, ...).

Every module here re-exports the implementation living in `bayesformer_b200/`, so the reference's
notebooks run unchanged against the B200-native path.
"""
