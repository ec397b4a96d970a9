from .CMixtureModel import CMixtureModel
from .CGaussianMixtureModel import CGaussianMixtureModel
