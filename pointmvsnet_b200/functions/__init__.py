"""Operator mirrors: dgcnn_ext shim, gather_knn autograd function, pixel grids."""
