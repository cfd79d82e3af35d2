from .group import ShardGroup

__all__ = ["ShardGroup"]
