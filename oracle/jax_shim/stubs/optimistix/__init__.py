class Newton: pass
