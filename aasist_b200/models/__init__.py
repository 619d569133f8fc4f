"""Plug-in modules named like the reference's ``models/`` package so that the reference's
``import_module("models.{architecture}")`` (main.py:253) can be pointed at
``aasist_b200.models.{architecture}`` unchanged."""
