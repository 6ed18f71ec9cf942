"""Command-line / config-file overrides for the driver scripts, with the semantics of the reference's
nanoGPT/configurator.py:20-47 (config files are Python that assigns globals; `--key=value` overrides an existing key and
must keep its type; unknown keys are errors) — implemented as a function instead of an exec'd script.

    settings = load_settings(defaults_dict, sys.argv[1:])
"""
from __future__ import annotations

from ast import literal_eval


def load_settings(defaults: dict, argv: list[str]) -> dict:
    settings = dict(defaults)
    for arg in argv:
        if "=" not in arg:
            if arg.startswith("--"):
                raise ValueError(f"expected --key=value, got {arg}")
            print(f"Overriding config with {arg}:")
            with open(arg) as f:
                source = f.read()
            print(source)
            scope: dict = {}
            exec(compile(source, arg, "exec"), {}, scope)  # a config file is plain assignments
            for key, val in scope.items():
                if key.startswith("_"):
                    continue
                settings[key] = val  # config files may introduce derived names; only known keys are read later
        else:
            if not arg.startswith("--"):
                raise ValueError(f"expected --key=value, got {arg}")
            key, val = arg[2:].split("=", 1)
            if key not in settings:
                raise ValueError(f"Unknown config key: {key}")
            try:
                parsed = literal_eval(val)
            except (SyntaxError, ValueError):
                parsed = val
            if type(parsed) is not type(settings[key]):
                raise TypeError(f"--{key}: expected {type(settings[key]).__name__}, got {type(parsed).__name__}")
            print(f"Overriding: {key} = {parsed}")
            settings[key] = parsed
    return settings
