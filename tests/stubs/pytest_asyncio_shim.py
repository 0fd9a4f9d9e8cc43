"""Test-only: runs `async def` tests with asyncio.run, for the reference's own router unit tests
(`@pytest.mark.asyncio`; the pytest-asyncio wheel is absent offline).  Loaded with `-p pytest_asyncio_shim`."""
import asyncio
import inspect

import pytest


def pytest_configure(config):
    config.addinivalue_line("markers", "asyncio: run the coroutine test in an event loop")


@pytest.hookimpl(tryfirst=True)
def pytest_pyfunc_call(pyfuncitem):
    if inspect.iscoroutinefunction(pyfuncitem.obj):
        args = {a: pyfuncitem.funcargs[a] for a in pyfuncitem._fixtureinfo.argnames}
        asyncio.run(pyfuncitem.obj(**args))
        return True
    return None
