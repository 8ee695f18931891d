from scipy.io import loadmat  # kernels_12.mat is MATLAB v5
