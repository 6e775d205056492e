"""Dataset constants consumed by the hot path (mirrors /root/reference/src/constants.py:11-54).

Only what the model / predictor need: the ten mice, their neuron counts, folds."""
new_mice = [
    "dynamic29515-10-12-Video-9b4f6a1a067fe51e15306b9628efea20",
    "dynamic29623-4-9-Video-9b4f6a1a067fe51e15306b9628efea20",
    "dynamic29647-19-8-Video-9b4f6a1a067fe51e15306b9628efea20",
    "dynamic29712-5-9-Video-9b4f6a1a067fe51e15306b9628efea20",
    "dynamic29755-2-8-Video-9b4f6a1a067fe51e15306b9628efea20",
]
old_mice = [
    "dynamic29156-11-10-Video-8744edeac3b4d1ce16b680916b5267ce",
    "dynamic29228-2-10-Video-8744edeac3b4d1ce16b680916b5267ce",
    "dynamic29234-6-9-Video-8744edeac3b4d1ce16b680916b5267ce",
    "dynamic29513-3-5-Video-8744edeac3b4d1ce16b680916b5267ce",
    "dynamic29514-2-9-Video-8744edeac3b4d1ce16b680916b5267ce",
]
new_num_neurons = [7863, 7908, 8202, 7939, 8122]   # constants.py:18
old_num_neurons = [7440, 7928, 8285, 7671, 7495]   # constants.py:26
mice = new_mice + old_mice
num_neurons = new_num_neurons + old_num_neurons    # constants.py:38
num_mice = len(mice)
index2mouse = dict(enumerate(mice))
mouse2index = {m: i for i, m in enumerate(mice)}
mouse2num_neurons = dict(zip(mice, num_neurons))
mice_indexes = list(range(num_mice))
num_folds = 7
folds = list(range(num_folds))
