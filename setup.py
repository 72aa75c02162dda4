import os

from setuptools import find_packages, setup

here = os.path.abspath(os.path.dirname(__file__))
version = {}
with open(os.path.join(here, 'patchgan_b200/version.py')) as f:
    exec(f.read(), version)

setup(
    name='patchgan_b200',
    version=version['__version__'],
    description='B200-native (sm_100a) implementation of the patchGAN training / inference hot path',
    packages=find_packages(include=['patchgan_b200', 'patchgan_b200.*']),
    package_data={'patchgan_b200': ['libpatchgan_b200.so', 'csrc/*']},
    entry_points={'console_scripts': ['patchgan_train = patchgan_b200.train:patchgan_train',
                                      'patchgan_infer = patchgan_b200.infer:patchgan_infer']},
    install_requires=['numpy', 'torch', 'tqdm', 'pyyaml'],
)
