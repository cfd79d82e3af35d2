from .group import LocalGroup, ShardGroup

__all__ = ["ShardGroup", "LocalGroup"]
