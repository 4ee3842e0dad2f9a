"""keds_b200: B200-native (sm_100a) knowledge-retrieval hot path of KEDs behind the reference's
own Faiss-shaped operator interface. Native code: keds_b200/csrc -> libkeds_knn.so (C ABI in
include/keds_knn.h). No CPU path."""
__version__ = "0.1.0"
