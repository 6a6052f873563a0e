MetadataLike = object
