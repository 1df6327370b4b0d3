"""bayesnf_b200: B200-native (sm_100a) BayesNF ensemble training / prediction.

Drop-in for the hot path of google/bayesnf: same estimator classes as
``bayesnf/__init__.py:21-23``.
"""

from .spatiotemporal import BayesianNeuralFieldMAP
from .spatiotemporal import BayesianNeuralFieldMLE
from .spatiotemporal import BayesianNeuralFieldVI

__all__ = ['BayesianNeuralFieldMAP', 'BayesianNeuralFieldMLE', 'BayesianNeuralFieldVI']
__version__ = '0.1.0'
