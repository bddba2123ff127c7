from .classifier import ClassifierGuidedDDIM, ClassifierGuidedDDPM, ClassifierMixin

__all__ = ["ClassifierMixin", "ClassifierGuidedDDPM", "ClassifierGuidedDDIM"]
