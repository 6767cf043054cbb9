"""Import alias: the product package lives in ``embodied-clip_b200/`` (the layout name the build
contract asks for, which is not a valid Python identifier).  ``import embclip_b200`` resolves its
sub-modules from that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "embodied-clip_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
