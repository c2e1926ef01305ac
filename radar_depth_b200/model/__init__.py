"""Mirror of the reference's ``model`` package (model/models.py, model/multistage_model.py) on the B200 kernels."""
