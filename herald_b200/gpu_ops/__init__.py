from .EmbeddingLookUp import (EmbeddingLookUp, EmbeddingLookUp_Gradient, embedding_lookup_op,
                              embedding_lookup_gradient_op)
from .ParameterServerCommunicate import ParameterServerCommunicateOp, parameterServerCommunicate_op

__all__ = ["EmbeddingLookUp", "EmbeddingLookUp_Gradient", "embedding_lookup_op",
           "embedding_lookup_gradient_op", "ParameterServerCommunicateOp",
           "parameterServerCommunicate_op"]
