"""Normal-inverse-Wishart / normal-inverse-chi-squared prior holder (segmentalist/niw.py:7-15)."""


class NIW(object):
    """A normal-inverse-Wishart distribution; with a vector S_0 it parametrises the D independent
    normal-inverse-chi-squared priors of the diagonal-covariance components."""

    def __init__(self, m_0, k_0, v_0, S_0):
        self.m_0 = m_0
        self.k_0 = k_0
        D = len(m_0)
        assert v_0 >= D, "v_0 must be larger or equal to dimension of data"
        self.v_0 = v_0
        self.S_0 = S_0
