class FlowMatchEulerDiscreteScheduler:   # annotation only (pipeline_chronoedit.py:168)
    pass
