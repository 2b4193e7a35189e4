class SVD: pass
